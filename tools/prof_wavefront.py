#!/usr/bin/env python3
"""One reconstruction of the 4K config-3 picture (all Main tools) or of a 4K I picture, for ncu captures of the wavefront kernel:
    ncu --set full --import-source on -k regex:k_recon_intra -c 1 -o out python tools/prof_wavefront.py [main|ibase|ieipd]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from xevd_b200 import synth  # noqa: E402
from xevd_b200.device import Context  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "main"
w, h, bd = 3840, 2160, 10
ctx = Context(0)
if kind == "main":
    prm, cl, refs, alf, flags = synth.make_main_frame(w, h, bit_depth=bd, seed=3)
    r1 = refs[::-1]
else:
    eipd = kind == "ieipd"
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=9, n_refs=1, coded_frac=0.7, iqt=eipd)
    prm.tool_eipd = prm.tool_htdf = int(eipd)
    prm.slice_qp = 34
    synth.add_intra_cus(cl, np.random.default_rng(2), 1.0, eipd=eipd)
    synth.derive_avail_cu(cl)
    refs = synth.make_refs(w, h, bd, 1, seed=7)
    r1 = []
d0 = [ctx.pic_alloc(w, h).upload(r) for r in refs]
d1 = d0[::-1] if r1 else []
cur = ctx.pic_alloc(w, h)
for _ in range(2):
    ctx.recon_frame(prm, cur, d0, d1, cl)
ctx.sync()
print("done", kind, cl.n_cu, "CUs")
