#!/usr/bin/env python3
"""A/B timing of the picture-wide passes (deblocking, ALF, padding) of two builds of libxevd_b200.so on one GPU in one process:
    python tools/ab_filters.py scratch/libxevd_b200_base.so xevd_b200/libxevd_b200.so
4K 10-bit pictures reconstructed from a Main-partition work list (non-square CUs, maps published by the recon kernels), 6 picture
slots in rotation (> L2), CUDA events on the launching stream, median; the builds alternate pass by pass.  Before timing, every build's
ALF and deblocking output on a small picture is compared with the CPU oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle  # noqa: E402
from xevd_b200 import abi, synth  # noqa: E402
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

# ---- parity (ALF: mirrored borders, partial tiles, CTU flags; deblocking: both filters on Main partitions) --------------------------
o = Oracle()
for (w, h, bd, log2_ctu, enable) in [(200, 136, 10, 6, (1, 1, 1)), (72, 200, 8, 5, (1, 0, 1)), (320, 192, 10, 7, (0, 1, 1))]:
    rng = np.random.default_rng(w + h)
    pic = HostPicture.random(w, h, bd, rng)
    prm = abi.make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, enable)
    n_ctu = ((w + (1 << log2_ctu) - 1) >> log2_ctu) * ((h + (1 << log2_ctu) - 1) >> log2_ctu)
    flags = (rng.random(n_ctu) < 0.7).astype(np.uint8)
    want = o.alf_frame(prm, pic.copy(), alf, flags)
    for p, c in zip(libs, ctxs):
        d = c.pic_alloc(w, h).upload(pic, padded=False)
        c.alf(prm, d, alf, flags)
        got = d.download()
        bad = [int((a != b).sum()) for a, b in zip(got.planes(), want.planes())]
        print(f"parity alf {w}x{h} bd{bd} ctu{1 << log2_ctu} {enable} {p}: {'ok' if not any(bad) else bad}", flush=True)
        d.free()
for addb in (0, 1):
    w, h, bd = 320, 192, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=5, n_refs=2, coded_frac=0.7, iqt=True, main_mv=True)
    prm.tool_addb = addb
    refs = synth.make_refs(w, h, bd, 2, seed=6)
    tbl = synth.chroma_qp_table(True)
    want = o.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    o.deblock_frame(prm, want, cl, tbl, bool(addb), ((0, 1), (1, 0)))
    for p, c in zip(libs, ctxs):
        dr = [c.pic_alloc(w, h).upload(r) for r in refs]
        cur = c.pic_alloc(w, h)
        c.set_chroma_qp_table(tbl)
        c.recon_frame(prm, cur, dr, dr[::-1], cl)
        c.deblock(prm, cur, dr, dr[::-1])
        got = cur.download()
        bad = [int((a != b).sum()) for a, b in zip(got.planes(), want.planes())]
        print(f"parity deblock addb={addb} {p}: {'ok' if not any(bad) else bad}", flush=True)
        for x in dr + [cur]:
            x.free()

# ---- timing -------------------------------------------------------------------------------------------------------------------------
w, h, bd, NPIC, REPS = 3840, 2160, 10, 6, 7
prm_m, cl_m = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=3, n_refs=2, coded_frac=0.7, iqt=True, main_mv=True)
host_refs = synth.make_refs(w, h, bd, 2, seed=7)
alf = synth.make_alf_params(np.random.default_rng(4))
state = []
for c in ctxs:
    drefs = [c.pic_alloc(w, h).upload(r) for r in host_refs]
    curs = [c.pic_alloc(w, h) for _ in range(NPIC)]
    c.set_chroma_qp_table(synth.chroma_qp_table(True))
    for cur in curs:
        c.recon_frame(prm_m, cur, drefs, drefs[::-1], cl_m)
    c.sync()
    state.append((drefs, curs))


def timed(fn):
    ts = []
    for _ in range(REPS):
        for i in range(NPIC):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(i); e1.record(stream)
            ts.append((e0, e1))
    torch.cuda.synchronize()
    return 1e3 * float(np.median([a.elapsed_time(b) for a, b in ts]))


def passes(c, drefs, curs):
    pb = type(prm_m).from_buffer_copy(prm_m); pb.tool_addb = 0
    pa = type(prm_m).from_buffer_copy(prm_m); pa.tool_addb = 1
    return [("deblock baseline", lambda i: c.deblock(pb, curs[i], drefs, drefs[::-1])),
            ("deblock addb", lambda i: c.deblock(pa, curs[i], drefs, drefs[::-1])),
            ("alf", lambda i: c.alf(prm_m, curs[i], alf, None)),
            ("pad", lambda i: c.pad(curs[i]))]


plist = [passes(c, *st) for c, st in zip(ctxs, state)]
for k in range(len(plist[0])):
    res = {p: [] for p in libs}
    for rnd in range(3):
        for p, pl in zip(libs, plist):
            name, fn = pl[k]
            for i in range(NPIC):
                fn(i)
            torch.cuda.synchronize()
            res[p].append(timed(fn))
    print(plist[0][k][0], {p: [round(x, 1) for x in v] for p, v in res.items()}, flush=True)
