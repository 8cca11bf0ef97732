#!/usr/bin/env python3
"""A/B timing of two builds of libxevd_b200.so on one GPU in one process (the method behind every kernel change in profiles/):
    python tools/ab_v2.py scratch/libxevd_b200_base.so xevd_b200/libxevd_b200.so
For every build: a parity check of small pictures against the CPU oracle (partitions down to 4x4, bi-prediction, IQT), then the 4K
workloads timed alternately A, B, A, B ... (CUDA events on the launching stream, 6 picture slots in rotation, median)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle  # noqa: E402
from xevd_b200 import synth  # noqa: E402
from xevd_b200.device import Context  # noqa: E402
from xevd_b200.frame import HostPicture  # noqa: E402

libs = sys.argv[1:]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctxs = []
for p in libs:
    c = Context(0, lib_path=str(ROOT / p))
    c.set_stream(stream.cuda_stream)
    ctxs.append(c)

# ---- parity ---------------------------------------------------------------------------------------------------------------------
o = Oracle()
cases = [("A16", dict(variant="A"), 1), ("A4", dict(variant="A", log2_cu=2), 1), ("A8bi", dict(variant="A", log2_cu=3, bi_frac=1.0), 2), ("B", dict(variant="B"), 2),
         ("C", dict(variant="C"), 2), ("C4", dict(variant="C", min_log2=2), 2), ("Biqt", dict(variant="B", iqt=True, main_mv=True), 2),
         ("C4iqt", dict(variant="C", min_log2=2, iqt=True, main_mv=True), 2), ("A64", dict(variant="A", log2_cu=6), 1), ("A32", dict(variant="A", log2_cu=5), 1)]
for bd in (8, 10):
    for name, kw, nl in cases:
        w, h = 320, 192
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, seed=3, n_refs=2 if nl == 2 else 1, coded_frac=0.8, **kw)
        refs = synth.make_refs(w, h, bd, 2, seed=4)
        r0, r1 = (refs[:1], []) if nl == 1 else (refs, refs[::-1])
        want = o.recon_frame(prm, HostPicture(w, h, prm.poc), r0, r1, cl)
        for p, c in zip(libs, ctxs):
            d0 = [c.pic_alloc(w, h).upload(r) for r in r0]
            d1 = d0[::-1] if nl == 2 else []
            cur = c.pic_alloc(w, h)
            c.recon_frame(prm, cur, d0, d1, cl)
            got = cur.download(maps=True)
            bad = [n for a, b, n in zip(got.planes(), want.planes(), "YUV") if not np.array_equal(a, b)]
            bad += [n for n in ("map_scu", "map_mv", "map_refi") if not np.array_equal(getattr(got, n), getattr(want, n))]
            if bad:
                print(f"PARITY FAIL {p} {name} {bd}-bit: {bad}")
            for x in d0 + [cur]:
                x.free()
print("parity checked", flush=True)

# ---- timing ---------------------------------------------------------------------------------------------------------------------
w, h, bd = 3840, 2160, 10
refs = synth.make_refs(w, h, bd, 2, seed=7)


def up(cl):
    return dict(cl=cl, cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev), first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev),
                ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev), coef=torch.from_numpy(cl.coef.copy()).to(dev), max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max()))


state = []
for c in ctxs:
    state.append(dict(refs=[c.pic_alloc(w, h).upload(r) for r in refs], curs=[c.pic_alloc(w, h) for _ in range(6)]))
work = [("2A uni 16x16", dict(variant="A"), 1), ("2B quadtree 50% bi", dict(variant="B"), 2), ("uni 8x8", dict(variant="A", log2_cu=3), 1),
        ("uni 32x32", dict(variant="A", log2_cu=5), 1), ("uni 64x64", dict(variant="A", log2_cu=6), 1), ("IQT quadtree", dict(variant="B", iqt=True, main_mv=True), 2),
        ("BTT C", dict(variant="C", main_mv=True), 2)]
for name, kw, nl in work:
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, seed=1, n_refs=2 if nl == 2 else 1, **kw)
    wk = up(cl)
    res = [[] for _ in ctxs]
    for rep in range(5):
        for k, c in enumerate(ctxs):
            st = state[k]
            r0, r1 = (st["refs"][:1], []) if nl == 1 else (st["refs"], st["refs"][::-1])

            def fn(i):
                c.recon_frame_dev(prm, st["curs"][i], r0, r1, wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(), cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext),
                                  wk["coef"].data_ptr(), cl.coef.size, has_intra=False, max_cu_per_ctu=wk["max_cu"])
            if rep == 0:
                for i in range(6):
                    fn(i)
                c.sync()
            ev = []
            for i in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(i); e1.record(stream)
                ev.append((e0, e1))
            c.sync()
            res[k] += [a.elapsed_time(b) for a, b in ev]
    print(f"{name:22s}" + "  ".join(f"{Path(p).name}: {1e3 * float(np.median(r)):7.1f} us" for p, r in zip(libs, res)), flush=True)
