#!/usr/bin/env python3
"""A/B timing of the CTU wavefront kernel (k_recon_intra) for two builds of libxevd_b200.so in one process: 4K I pictures (Baseline modes;
EIPD + HTDF) and a Main P picture with HTDF on 60 % of the CUs, device-resident work lists, 6 picture slots in rotation.
    python tools/ab_wave.py scratch/libxevd_b200_base.so xevd_b200/libxevd_b200.so"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from xevd_b200 import synth  # noqa: E402
from xevd_b200.device import Context  # noqa: E402

libs = sys.argv[1:]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
w, h, bd, NPIC = 3840, 2160, 10, 4
host_refs = synth.make_refs(w, h, bd, 2, seed=7)
ctxs = []
for p in libs:
    c = Context(0, lib_path=str(ROOT / p))
    c.set_stream(stream.cuda_stream)
    ctxs.append((c, [c.pic_alloc(w, h).upload(r) for r in host_refs], [c.pic_alloc(w, h) for _ in range(NPIC)]))


def cases():
    for eipd in (0, 1):
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=9, n_refs=1, coded_frac=0.7, iqt=bool(eipd))
        prm.tool_eipd = eipd; prm.tool_htdf = eipd; prm.slice_qp = 34
        synth.add_intra_cus(cl, np.random.default_rng(2), 1.0, eipd=bool(eipd))
        synth.derive_avail_cu(cl)
        yield ("I picture, " + ("EIPD + HTDF" if eipd else "Baseline modes"), prm, cl, 1)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=23, n_refs=2, coded_frac=0.6, iqt=True, main_mv=True)
    prm.tool_eipd = prm.tool_htdf = 1; prm.slice_qp = 34
    synth.add_intra_cus(cl, np.random.default_rng(7), 0.02, eipd=True)
    synth.derive_avail_cu(cl)
    yield ("Main P picture, HTDF on 60 % of the CUs", prm, cl, 2)
    prm, cl, refs, _, _ = synth.make_main_frame(w, h, bit_depth=bd, seed=3)
    yield ("config 3 picture (all Main tools: throughput + generic + wavefront kernel)", prm, cl, refs)


for name, prm, cl, nl in cases():
    wk = dict(cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev), first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev),
              ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev), coef=torch.from_numpy(cl.coef.copy()).to(dev),
              max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max()))
    res, sums = [], []
    for c, drefs, curs in ctxs:
        own = None
        if not isinstance(nl, int):         # the case brings its own reference pictures
            own = [c.pic_alloc(w, h).upload(r) for r in nl]
            drefs = own
        nl_ = 2 if own else nl
        def run(i, nl=nl_, drefs=drefs):
            c.recon_frame_dev(prm, curs[i], drefs[:nl] if nl == 1 else drefs, [] if nl == 1 else drefs[::-1], wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(),
                              cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext), wk["coef"].data_ptr(), cl.coef.size, has_intra=True, max_cu_per_ctu=wk["max_cu"])
        for i in range(NPIC):
            run(i)
        c.sync()
        ts = []
        for rep in range(3):
            for i in range(NPIC):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); run(i); e1.record(stream)
                ts.append((e0, e1))
        torch.cuda.synchronize()
        res.append(1e3 * float(np.median([a.elapsed_time(b) for a, b in ts])))
        got = curs[0].download()
        sums.append(int(got.y.astype(np.int64).sum() + 3 * got.u.astype(np.int64).sum() + 7 * got.v.astype(np.int64).sum()))
        for p_ in own or []:
            p_.free()
    print(f"{name}: " + "  ".join(f"{Path(p).name} {r:.0f} us" for p, r in zip(libs, res)) + f"  identical: {len(set(sums)) == 1}", flush=True)
