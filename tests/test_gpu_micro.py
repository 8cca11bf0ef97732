"""GPU parity of the batched leaf entry points (BASELINE config 5): xb200_itdq_blocks_dev and xb200_mc_blocks_dev through the C ABI
against the CPU oracle and against the golden vectors recorded from the unmodified reference (tests/golden/{itdq,mc}_blocks.npz).

  xevd_itdq / xevdm_itdq       src_base/xevd_itdq.c:494-542, src_main/xevdm_itdq.c:708-788   (dequant + 2-D inverse DCT-2, all 36 shapes)
  xevd_mc_l / xevd_mc_c        src_base/xevd_mc.h:66-74 -> xevd_mc.c:169-408                 (00 / n0 / 0n / nn, T3 variant-vs-phase split)
"""
from pathlib import Path

import numpy as np
import pytest

from xevd_b200 import synth
from xevd_b200.frame import HostPicture

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ctx():
    from xevd_b200.device import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch_dev():
    import torch
    return torch, torch.device("cuda", 0)


def _gpu_itdq(ctx, torch_dev, blocks, qp, bd, iqt):
    """blocks: (n, h, w) int16 levels -> (n, h, w) int16 residual through xb200_itdq_blocks_dev"""
    torch, dev = torch_dev
    n, h, w = blocks.shape
    d_in = torch.from_numpy(np.ascontiguousarray(blocks).reshape(-1).copy()).to(dev)
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    ctx.itdq_blocks_dev(d_in.data_ptr(), d_out.data_ptr(), n, int(np.log2(w)), int(np.log2(h)), qp, bd, bool(iqt))
    ctx.sync()
    return d_out.cpu().numpy().reshape(n, h, w)


@pytest.mark.parametrize("iqt", [0, 1])
@pytest.mark.parametrize("bd", [8, 10])
def test_itdq_blocks_golden(ctx, torch_dev, iqt, bd):
    """every (log2w, log2h) in 1..6 x 1..6 against the reference's recorded output"""
    z = np.load(G / "itdq_blocks.npz")
    qp = 32 + 6 * (bd - 8)
    n = 0
    for lw in range(1, 7):
        for lh in range(1, 7):
            key = f"{iqt}_{bd}_{lw}_{lh}"
            got = _gpu_itdq(ctx, torch_dev, z["in_" + key][None], qp, bd, iqt)[0]
            assert np.array_equal(got, z["out_" + key]), key
            n += 1
    assert n == 36


@pytest.mark.parametrize("iqt", [0, 1])
@pytest.mark.parametrize("bd,qp", [(8, 27), (10, 44), (10, 39), (12, 58)])
def test_itdq_blocks_vs_oracle(ctx, oracle, torch_dev, iqt, bd, qp):
    """batches of blocks (ragged last CTA group) of every shape against the oracle, several QPs (all six qp % 6 scales occur over the cases)"""
    rng = np.random.default_rng(100 * bd + qp + iqt)
    for lw in range(1, 7):
        for lh in range(1, 7):
            nb = int(rng.integers(3, 12))
            res = rng.laplace(0, 30.0 * (1 << (bd - 8)), (nb, 1 << lh, 1 << lw))
            lev = np.stack([synth.quantised_dct(r, qp, bool(iqt)) for r in res])
            if iqt and lw == 6:
                lev[:, :, 32:] = 0          # conforming IQT streams carry zeros there
            if iqt and lh == 6:
                lev[:, 32:, :] = 0
            got = _gpu_itdq(ctx, torch_dev, lev, qp, bd, iqt)
            for b in range(nb):
                want = oracle.itdq_block(lev[b], qp, bd, iqt)
                assert np.array_equal(got[b], want), (iqt, bd, qp, lw, lh, b)


def _gpu_mc(ctx, torch_dev, dpic, plane, mvs, w, h, bd, main):
    torch, dev = torch_dev
    n = len(mvs)
    d_mv = torch.from_numpy(np.ascontiguousarray(mvs, np.int32)).to(dev)
    d_out = torch.zeros(n * w * h, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    ctx.mc_blocks_dev(dpic, plane, d_mv.data_ptr(), d_out.data_ptr(), n, w, h, bd, bool(main))
    ctx.sync()
    return d_out.cpu().numpy().reshape(n, h, w)


@pytest.mark.parametrize("bd", [8, 10])
def test_mc_blocks_golden(ctx, torch_dev, bd):
    """the 48 recorded cases per bit depth (luma / chroma, Baseline / Main taps, T3 cases where the variant and the phase disagree)"""
    z = np.load(G / "mc_blocks.npz")
    plane = z[f"plane_{bd}"]                          # (96, 112); the recorded calls use sample (24, 20) as origin
    ph, pw = plane.shape
    # luma cases: the plane is the luma interior of a 112x96 picture; chroma cases: it is the Cb interior of a 224x192 picture
    pl = HostPicture(pw, ph, 0); pl.y[...] = plane
    pc = HostPicture(2 * pw, 2 * ph, 0); pc.u[...] = plane
    dl = ctx.pic_alloc(pw, ph).upload(pl)
    dc = ctx.pic_alloc(2 * pw, 2 * ph).upload(pc)
    for t, (w, h, gx, gy, ox, oy, chroma, main) in enumerate(z[f"cases_{bd}"]):
        sh = 5 if chroma else 4
        mv = [[int(gx) + (24 << sh), int(gy) + (20 << sh), int(ox), int(oy)]]
        got = _gpu_mc(ctx, torch_dev, dc if chroma else dl, 1 if chroma else 0, mv, int(w), int(h), bd, main)[0]
        assert np.array_equal(got, z[f"mc_{bd}_{t}"]), (bd, t)
    dl.free(); dc.free()


@pytest.mark.parametrize("main", [0, 1])
@pytest.mark.parametrize("chroma", [0, 1])
@pytest.mark.parametrize("bd", [8, 10])
def test_mc_blocks_vs_oracle(ctx, oracle, torch_dev, bd, chroma, main):
    """sizes 4..128 (chroma 2..64), all four variants 00 / n0 / 0n / nn, vectors reaching into the padded border"""
    rng = np.random.default_rng(7 + bd + 2 * chroma + main)
    W, H = 256, 192
    pic = HostPicture.random(W, H, bd, rng)
    oracle.pad(pic)
    d = ctx.pic_alloc(W, H).upload(pic)
    plane = pic.buf_u if chroma else pic.buf_y
    pad = pic.pad_c if chroma else pic.pad_l
    pw, ph = (W >> 1, H >> 1) if chroma else (W, H)
    sh = 5 if chroma else 4
    nph = 1 << sh
    step = 1 if main else 4
    sizes = [2, 4, 8, 16, 32, 64] if chroma else [4, 8, 16, 32, 64, 128]
    for w in sizes:
        for h in sizes:
            if max(w, h) > 8 * min(w, h):
                continue
            mvs = []
            for k in range(6):
                variant = k & 3             # 0: 00, 1: n0, 2: 0n, 3: nn
                fx = int(rng.integers(1, nph // step)) * step if variant & 1 else 0
                fy = int(rng.integers(1, nph // step)) * step if variant & 2 else 0
                x = int(rng.integers(-60, pw - w + 60)); y = int(rng.integers(-60, ph - h + 60))
                ox, oy = fx, fy
                if k >= 4:                  # T3: variant chosen by the unclipped vector, phase by the clipped one
                    ox, oy = int(rng.integers(0, nph)), int(rng.integers(0, nph))
                mvs.append([(x << sh) + fx, (y << sh) + fy, ox, oy])
            got = _gpu_mc(ctx, torch_dev, d, 1 if chroma else 0, mvs, w, h, bd, main)
            for k, (gx, gy, ox, oy) in enumerate(mvs):
                want = oracle.mc(plane, (pad, pad), (gx, gy), (ox, oy), w, h, bd, bool(chroma), bool(main))
                assert np.array_equal(got[k], want), (bd, chroma, main, w, h, k)
    d.free()
