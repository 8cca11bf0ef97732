#!/usr/bin/env python3
"""One 4K Main-partition picture through deblocking (both filters), ALF and padding - the command the ncu captures of the picture-wide
passes are taken from:
    ncu --set full --clock-control none --import-source on -k regex:'k_alf|k_deblock|k_pad' -o gpurun_out/filters python tools/prof_filters.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from xevd_b200 import synth  # noqa: E402
from xevd_b200.device import Context  # noqa: E402

w, h, bd = 3840, 2160, 10
prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=3, n_refs=2, coded_frac=0.7, iqt=True, main_mv=True)
refs = synth.make_refs(w, h, bd, 2, seed=7)
alf = synth.make_alf_params(np.random.default_rng(4))
with Context(0) as c:
    dr = [c.pic_alloc(w, h).upload(r) for r in refs]
    c.set_chroma_qp_table(synth.chroma_qp_table(True))
    for addb in (0, 1):
        cur = c.pic_alloc(w, h)
        c.recon_frame(prm, cur, dr, dr[::-1], cl)
        p2 = type(prm).from_buffer_copy(prm)
        p2.tool_addb = addb
        c.deblock(p2, cur, dr, dr[::-1])
        c.alf(prm, cur, alf, None)
        c.pad(cur)
        c.sync()
        cur.free()
