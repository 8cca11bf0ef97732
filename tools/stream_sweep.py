#!/usr/bin/env python3
"""Random-configuration sweep of the drop-in decoder against the unmodified reference.
    here (needs /root/reference):  python tools/stream_sweep.py gen scratch/sweep 24      # tools/evcgen draws 24 small streams with random
                                                                                          # profile / size / CTU / tile grid / slices / tools
    on a B200:                     python tools/stream_sweep.py check scratch/sweep      # libxevd_gpu.so vs libxevd_ref.so, picture by picture
Streams the reference does not reproduce itself (generator vs reference) are not kept."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools" / "evcgen"))
from xevd_b200 import xevd_api as X  # noqa: E402


def gen(out, n, seed0=1000, skip=0, per_stream_s=240):
    import json
    import subprocess
    import evcgen as G
    out.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(seed0)
    kept = 0
    for k in range(n):
        main = bool(rng.integers(0, 2))
        tools = dict(G.MAIN if main else G.BASELINE)
        log2_ctu = int(rng.choice([5, 6, 7])) if main else 6
        ctu = 1 << log2_ctu
        w = int(rng.integers(2, 7)) * ctu // 2 // 8 * 8 + int(rng.choice([0, 8, 24]))
        h = int(rng.integers(2, 5)) * ctu // 2 // 8 * 8 + int(rng.choice([0, 8, 16]))
        wc, hc = (w + ctu - 1) // ctu, (h + ctu - 1) // ctu
        cols, rows = int(rng.integers(1, min(wc, 3) + 1)), int(rng.integers(1, min(hc, 3) + 1))
        tiles = None if cols * rows == 1 else dict(cols=cols, rows=rows, across=int(rng.integers(0, 2)))
        if tiles and rng.random() < 0.4 and cols > 1:
            cw = [1] * cols; cw[int(rng.integers(0, cols))] += wc - cols
            rh = [1] * rows; rh[int(rng.integers(0, rows))] += hc - rows
            tiles.update(col_w=cw, row_h=rh)
        idr_only = bool(tiles) and rows > 1 and rng.random() < 0.4
        slices = [(0, cols - 1), (cols, cols * rows - 1)] if idr_only else None     # first tile row | the rest
        if main:
            tools.update(alf=int(rng.integers(0, 2)), ibc=0 if tiles else int(rng.integers(0, 2)), dmvr=int(rng.integers(0, 2)),
                         affine=int(rng.integers(0, 2)), addb=int(rng.integers(0, 2)), htdf=int(rng.integers(0, 2)))
        kw = dict(w=w, h=h, bd=int(rng.choice([8, 10])), frames=int(rng.integers(2, 5)), seed=seed0 + 17 * k, types="I" if idr_only else str(rng.choice(["IPB", "IBB", "IPP"])),
                  qp=int(rng.integers(24, 40)), lps_scale=350 if main else 256, log2_ctu=log2_ctu, tiles=tiles, slices=slices, gop=1 if idr_only else 0)
        if k < skip:
            continue
        name = f"sweep_{k:03d}_{'main' if main else 'base'}_{w}x{h}_ctu{ctu}" + (f"_t{cols}x{rows}{'a' if tiles['across'] else 'n'}" if tiles else "") + ("_sl2" if slices else "")
        # one process per stream: a draw the generator cannot finish (it can loop inside the reference's parser) is cut off by the timeout
        try:
            r = subprocess.run([sys.executable, __file__, "one", str(out / f"{name}.evc"), json.dumps(dict(main=main, tools=tools, kw=kw))],
                               timeout=per_stream_s, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("one:")]
            print(f"{k}: {name}: {line[-1][5:] if line else 'failed: ' + r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else 'failed'}", flush=True)
            kept += (out / f"{name}.evc").exists()
        except subprocess.TimeoutExpired:
            print(f"{k}: {name}: generation did not finish in {per_stream_s} s, skipped (tools {tools}, {kw})", flush=True)
    print(f"{kept} streams kept in {out}")


def one(path, spec):
    """generate one stream (own process); keep it only if the unmodified reference reproduces the generator's pictures"""
    import json
    import evcgen as G
    spec = json.loads(spec)
    kw = spec["kw"]
    if kw.get("slices"):
        kw["slices"] = [tuple(x) for x in kw["slices"]]
    try:
        nals, own = G.Generator().make(spec["tools"], **kw)
    except (G.NonConforming, AssertionError) as e:
        print(f"one: no stream ({str(e)[:80]})")
        return
    pics = X.decode_stream(X.XevdLibrary(X.REF_SO), nals)
    if not G.same_pictures(own, pics):
        print("one: reference != generator, dropped")
        return
    X.write_stream(path, nals)
    print(f"one: {len(pics)} pictures, {sum(map(len, nals))} bytes")


def check(d):
    gpu, ref = X.XevdLibrary(X.GPU_SO), X.XevdLibrary(X.REF_SO)
    bad = 0
    files = sorted(d.glob("*.evc"))
    for f in files:
        nals = X.read_stream(f)
        a, b = X.decode_stream(gpu, nals), X.decode_stream(ref, nals)
        diffs = [(i, [int((x != y).sum()) for x, y in zip(pa, pb)]) for i, (pa, pb) in enumerate(zip(a, b)) if not all(np.array_equal(x, y) for x, y in zip(pa, pb))]
        if len(a) != len(b) or diffs:
            bad += 1
            print(f"{f.stem}: MISMATCH {len(a)} vs {len(b)} pictures, {diffs[:3]}", flush=True)
        else:
            print(f"{f.stem}: {len(a)} pictures identical", flush=True)
    print(f"{len(files) - bad} of {len(files)} streams bit-identical")
    return 1 if bad else 0


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "gen":
        gen(Path(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 16, int(sys.argv[4]) if len(sys.argv) > 4 else 1000,
            int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    else:
        sys.exit(check(Path(sys.argv[2])))
