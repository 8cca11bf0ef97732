#!/usr/bin/env python3
"""A/B timing of xb200_mc_blocks_dev (BASELINE config 5: 8-tap luma / 4-tap chroma `nn` on 4 Mi samples per launch from a 4K plane) for two
builds of libxevd_b200.so in one process:   python tools/ab_mc.py scratch/libxevd_b200_base.so xevd_b200/libxevd_b200.so"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from xevd_b200.device import Context  # noqa: E402
from xevd_b200.frame import HostPicture  # noqa: E402

libs = sys.argv[1:]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
W, H, bd = 3840, 2160, 10
pic = HostPicture.random(W, H, bd, np.random.default_rng(5))
pic.pad_borders()
ctxs, pics = [], []
for p in libs:
    c = Context(0, lib_path=str(ROOT / p))
    c.set_stream(stream.cuda_stream)
    ctxs.append(c)
    pics.append(c.pic_alloc(W, H).upload(pic))
for chroma in (0, 1):
    for s in (8, 16, 32, 64):
        bw = s >> chroma
        pw, ph = W >> chroma, H >> chroma
        sh = 5 if chroma else 4
        nb = (1 << 22) // (bw * bw)
        rng = np.random.default_rng(s + chroma)
        x = rng.integers(-16, pw - bw + 16, nb); y = rng.integers(-16, ph - bw + 16, nb)
        fx = rng.integers(1, 4, nb) * (1 << (sh - 2)); fy = rng.integers(1, 4, nb) * (1 << (sh - 2))
        mv = np.stack([(x << sh) + fx, (y << sh) + fy, fx, fy], 1).astype(np.int32)
        d_mv = torch.from_numpy(mv).to(dev)
        outs = [torch.zeros(nb * bw * bw, dtype=torch.int16, device=dev) for _ in libs]
        res = []
        for c, dp, o in zip(ctxs, pics, outs):
            ts = []
            for k in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); c.mc_blocks_dev(dp, 1 if chroma else 0, d_mv.data_ptr(), o.data_ptr(), nb, bw, bw, bd, False); e1.record(stream)
                ts.append((e0, e1))
            torch.cuda.synchronize()
            res.append(1e3 * float(np.median([a.elapsed_time(b) for a, b in ts[2:]])))
        same = all(torch.equal(outs[0], o) for o in outs[1:])
        print(f"{'chroma' if chroma else 'luma'} {bw}x{bw}: " + "  ".join(f"{Path(p).name} {r:.1f} us" for p, r in zip(libs, res)) + f"  identical: {same}", flush=True)
