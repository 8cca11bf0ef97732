"""Synthetic pre-parsed CU arrays (BASELINE.json configs 2-5).

There are no EVC bitstreams on the build or GPU boxes, so workloads are generated the way SURVEY 8(d)
specifies: random reference pictures, a CU partition, quarter-pel motion vectors and coefficients that are
the quantised forward DCT of bounded residuals (so every inverse-transform sum stays in range, T4).

Deterministic: everything derives from the integer seed.
"""
from __future__ import annotations

import numpy as np

from .abi import CU_DTYPE, EXT_DTYPE, CUF_CHROMA, CUF_LUMA, MODE_INTER, MODE_INTRA, make_params
from .frame import CuList, HostPicture

_DQ_BASE = (40, 45, 51, 57, 64, 71)


def dq_scale(qp: int, iqt: bool = False) -> int:
    s = _DQ_BASE[qp % 6]
    if iqt and qp % 6 == 5:
        s = 72
    return s << (qp // 6)


def _orth_dct(n: int) -> np.ndarray:
    k = np.arange(n)[:, None]
    i = np.arange(n)[None, :]
    m = np.cos(np.pi * (2 * i + 1) * k / (2 * n)) * np.sqrt(2.0 / n)
    m[0, :] = np.sqrt(1.0 / n)
    return m


def quantised_dct(res: np.ndarray, qp: int, iqt: bool = False) -> np.ndarray:
    """res: [..., h, w] float residual -> int16 levels such that dequant + inverse DCT ~ res.
    One level step is worth dq_scale(qp)/64 in the sample domain (derivation in DESIGN.md)."""
    h, w = res.shape[-2:]
    c = _orth_dct(h) @ res @ _orth_dct(w).T
    lev = np.rint(c * 64.0 / dq_scale(qp, iqt))
    if iqt:
        # 64-point IQT kernels only see the 32 lowest frequencies: the reference's dispatched AVX2
        # xevdm_itx_pb64_avx (src_main/avx/xevdm_itdq_avx.c:1147) ignores coefficients 32..63, the plain-C
        # xevdm_itx_pb64 does not.  Conforming IQT streams carry zeros there, so generate zeros.
        if h == 64:
            lev[..., 32:, :] = 0
        if w == 64:
            lev[..., :, 32:] = 0
    return np.clip(lev, -32768, 32767).astype(np.int16)


def partition_uniform(w: int, h: int, log2_ctu: int, log2_cu: int):
    """uniform CU grid, decoding order = z-order inside each CTU, CTUs in raster order"""
    out = []
    first = [0]
    ctu = 1 << log2_ctu

    def rec(x, y, lg):
        if x >= w or y >= h:
            return
        s = 1 << lg
        if lg > log2_cu or x + s > w or y + s > h:
            hs = s >> 1
            for dy in (0, hs):
                for dx in (0, hs):
                    rec(x + dx, y + dy, lg - 1)
        else:
            out.append((x, y, lg, lg))

    for cy in range(0, h, ctu):
        for cx in range(0, w, ctu):
            rec(cx, cy, log2_ctu)
            first.append(len(out))
    return out, first


def partition_quadtree(w: int, h: int, log2_ctu: int, rng, leaf_prob=(0.1, 1.0 / 3.0, 2.0 / 3.0)):
    """random quadtree 64/32/16/8 with leaf fractions {.1,.3,.4,.2} by area (SURVEY 8d variant B)"""
    out = []
    first = [0]
    ctu = 1 << log2_ctu

    def rec(x, y, lg, depth):
        if x >= w or y >= h:
            return
        s = 1 << lg
        must_split = x + s > w or y + s > h
        can_split = lg > 3
        p = leaf_prob[depth] if depth < len(leaf_prob) else 1.0
        if can_split and (must_split or rng.random() >= p):
            hs = s >> 1
            for dy in (0, hs):
                for dx in (0, hs):
                    rec(x + dx, y + dy, lg - 1, depth + 1)
        else:
            out.append((x, y, lg, lg))

    for cy in range(0, h, ctu):
        for cx in range(0, w, ctu):
            rec(cx, cy, log2_ctu, 0)
            first.append(len(out))
    return out, first


def make_inter_frame(w: int, h: int, *, bit_depth: int = 10, variant: str = "A", seed: int = 1,
                     n_refs: int = 1, bi_frac: float | None = None, coded_frac: float = 1.0,
                     mv_range_px: int = 32, resid_scale: float | None = None, iqt: bool = False,
                     log2_cu: int = 4, main_mv: bool = False):
    """Config-2 style inter picture.

    variant "A": uniform 16x16 CUs, uni-prediction, every CU coded in all three planes.
    variant "B": random quadtree {64,32,16,8}, 50 % bi-prediction.
    Returns (params, CuList).  Reference pictures are produced separately (make_refs).
    """
    rng = np.random.default_rng(seed)
    log2_ctu = 6
    if variant == "A":
        parts, first = partition_uniform(w, h, log2_ctu, log2_cu)
        bi = 0.0 if bi_frac is None else bi_frac
    else:
        parts, first = partition_quadtree(w, h, log2_ctu, rng)
        bi = 0.5 if bi_frac is None else bi_frac
    n = len(parts)
    cus = np.zeros(n, CU_DTYPE)
    pa = np.array(parts, np.int64)
    cus["x"], cus["y"], cus["log2w"], cus["log2h"] = pa[:, 0], pa[:, 1], pa[:, 2], pa[:, 3]
    cus["mode"] = MODE_INTER
    cus["flags"] = CUF_LUMA | CUF_CHROMA
    qp_y = 32 + 6 * (bit_depth - 8)
    cus["qp_y"] = qp_y
    cus["qp_u"] = qp_y
    cus["qp_v"] = qp_y
    cus["qp_map"] = 32
    # motion: quarter-pel, uniform in +-mv_range_px (all four interpolation variants occur)
    mv = rng.integers(-mv_range_px * 4, mv_range_px * 4 + 1, (n, 2, 2)).astype(np.int16)
    is_bi = rng.random(n) < bi
    use_l1 = (~is_bi) & (rng.random(n) < (0.5 if bi > 0 else 0.0))
    refi = np.full((n, 2), -1, np.int8)
    r0 = rng.integers(0, n_refs, n).astype(np.int8)
    r1 = rng.integers(0, n_refs, n).astype(np.int8)
    refi[:, 0] = np.where(use_l1, -1, r0)
    refi[:, 1] = np.where(is_bi | use_l1, r1, -1)
    mv[refi[:, 0] < 0, 0, :] = 0
    mv[refi[:, 1] < 0, 1, :] = 0
    cus["refi"] = refi
    cus["mv"] = mv
    coded = rng.random(n) < coded_frac
    cus["cbf"] = np.where(coded, 0x111, 0).astype(np.uint16)

    # coefficient stream: quantised forward DCT of Laplacian residuals
    if resid_scale is None:
        # ~10 % non-zero levels: Laplacian scale ~ 0.22 of a quantiser step (dq_scale/64)
        resid_scale = 0.22 * dq_scale(qp_y, iqt) / 64.0
    sizes = 1 << (cus["log2w"].astype(np.int64) + cus["log2h"].astype(np.int64))
    a8 = lambda v: (v + 7) & ~7          # plane blocks padded to multiples of 8 int16
    per_cu = np.where(coded, a8(sizes) + 2 * a8(sizes // 4), 0)
    off = np.concatenate(([0], np.cumsum(per_cu)))
    cus["coef_off"] = off[:-1].astype(np.uint32)
    coef = np.zeros(int(off[-1]), np.int16)
    for lg in np.unique(cus["log2w"]):
        sel = np.nonzero((cus["log2w"] == lg) & coded)[0]
        if len(sel) == 0:
            continue
        s = 1 << int(lg)
        ry = rng.laplace(0.0, resid_scale, (len(sel), s, s))
        rc = rng.laplace(0.0, resid_scale, (len(sel), 2, s // 2, s // 2))
        ly = quantised_dct(ry, qp_y, iqt).reshape(len(sel), -1)
        lc = quantised_dct(rc, qp_y, iqt).reshape(len(sel), 2, -1)
        nc = a8(s * s // 4)
        lcp = np.zeros((len(sel), 2, nc), np.int16)
        lcp[:, :, :lc.shape[2]] = lc
        blk = np.concatenate([ly, lcp.reshape(len(sel), -1)], axis=1)
        idx = off[sel][:, None] + np.arange(blk.shape[1])[None, :]
        coef[idx] = blk
    prm = make_params(w, h, bit_depth=bit_depth, log2_ctu=log2_ctu, poc=8, tool_iqt=int(iqt), tool_admvp=int(main_mv))
    cl = CuList(w=w, h=h, log2_ctu=log2_ctu, cus=cus, ctu_first=np.array(first, np.uint32), coef=coef)
    return prm, cl


def make_refs(w: int, h: int, bit_depth: int, n: int, seed: int = 1):
    """n random reference pictures with replicated borders; POCs 0, 16, 4, 12 ... around the current 8"""
    rng = np.random.default_rng(seed)
    pocs = [0, 16, 4, 12, 2, 14, 6, 10]
    return [HostPicture.random(w, h, bit_depth, rng, poc=pocs[i % len(pocs)]).pad_borders() for i in range(n)]


# default chroma QP mapping tables (xevd_tbl_qp_chroma_adjust_base / _main, src_base/xevd_tbl.c:334-355): what
# xevd_qp_chroma_dynamic holds when the SPS carries no chroma_qp_table
CHROMA_QP_BASE = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                           29, 29, 30, 31, 32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 39, 39, 40, 40, 40, 41, 41, 41], np.int32)
CHROMA_QP_MAIN = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                           29, 30, 31, 32, 33, 34, 35, 36, 37, 37, 38, 39, 40, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54], np.int32)


def chroma_qp_table(main: bool = False) -> np.ndarray:
    t = CHROMA_QP_MAIN if main else CHROMA_QP_BASE
    return np.stack([t, t]).copy()


def randomize_deblock_maps(pic: HostPicture, cl: CuList, rng, intra_frac=0.1, qp_lo=18, qp_hi=51):
    """publish the CUs' motion into map_mv / map_refi (what xevd_set_dec_info leaves there) and give every CU of a reconstructed picture its own QP / intra flag in map_scu (per-CU constant, as the decoder
    would publish them) so that all four deblocking strength classes and a wide QP range occur"""
    ws = pic.w_scu
    for cu in cl.cus:
        x0, y0 = int(cu["x"]) >> 2, int(cu["y"]) >> 2
        nw, nh = 1 << (int(cu["log2w"]) - 2), 1 << (int(cu["log2h"]) - 2)
        intra = int(cu["mode"]) == MODE_INTRA
        for j in range(nh):
            sl = slice((y0 + j) * ws + x0, (y0 + j) * ws + x0 + nw)
            pic.map_mv[sl] = 0 if intra else cu["mv"]
            pic.map_refi[sl] = -1 if intra else cu["refi"]
        qp = int(rng.integers(qp_lo, qp_hi + 1))
        m = (1 << 31) | (qp << 16) | (int(rng.random() < intra_frac) << 15) | ((int(cu["cbf"]) & 1) << 24)
        for j in range(nh):
            pic.map_scu[(y0 + j) * ws + x0:(y0 + j) * ws + x0 + nw] = m
    return pic


def add_intra_cus(cl: CuList, rng, intra_frac: float = 1.0, n_modes: int = 5, constrained: bool = False):
    """Turn a fraction of the CUs of a picture into intra CUs (Baseline modes 0..4 for luma and chroma) and derive their
    neighbour-availability masks the way the decoder does: a neighbouring SCU is available when it lies inside the picture
    and was reconstructed earlier in decoding order (COD bit of map_scu; xevd_get_nbr_b, src_base/xevd_ipred.c:49-91).
    With `constrained` (pps.constrained_intra_pred_flag) it must also be intra."""
    cus = cl.cus
    n = len(cus)
    w_scu, h_scu = (cl.w + 3) >> 2, (cl.h + 3) >> 2
    is_intra = rng.random(n) < intra_frac
    cus["mode"] = np.where(is_intra, MODE_INTRA, MODE_INTER)
    ipm = rng.integers(0, n_modes, (n, 2)).astype(np.int8)
    cus["refi"] = np.where(is_intra[:, None], ipm, cus["refi"])
    cus["mv"][is_intra] = 0
    ext = [np.zeros(1, EXT_DTYPE)[0]]
    cod = np.zeros((h_scu, w_scu), bool)
    intra_map = np.zeros((h_scu, w_scu), bool)
    for i in range(n):
        cu = cus[i]
        xs, ys = int(cu["x"]) >> 2, int(cu["y"]) >> 2
        nw, nh = 1 << (int(cu["log2w"]) - 2), 1 << (int(cu["log2h"]) - 2)
        if is_intra[i]:
            ok = (lambda yy, xx: cod[yy, xx] and (not constrained or intra_map[yy, xx]))
            up = left = 0
            for k in range(nw + nh):
                if ys > 0 and xs + k < w_scu and ok(ys - 1, xs + k):
                    up |= 1 << k
                if xs > 0 and ys + k < h_scu and ok(ys + k, xs - 1):
                    left |= 1 << k
            ul = int(ys > 0 and xs > 0 and ok(ys - 1, xs - 1))
            e = np.zeros(1, EXT_DTYPE)[0]
            e["q"][0], e["q"][1] = up, left
            cu["avail"] = (int(cu["avail"]) & 3) | (ul << 2)
            cu["mv"][1] = np.frombuffer(np.uint32(len(ext)).tobytes(), np.int16)
            ext.append(e)
        cod[ys:ys + nh, xs:xs + nw] = True
        intra_map[ys:ys + nh, xs:xs + nw] = is_intra[i]
    cl.ext = np.array(ext, EXT_DTYPE)
    return cl


def make_alf_params(rng, enable=(1, 1, 1)):
    """random but well-formed ALF filters: every filter sums to 512 (unity gain at shift 9), side taps within the ranges
    alf_recon_coef enforces (src_main/xevdm_alf.c:751,763)"""
    from .abi import AlfParams
    a = AlfParams()
    for c in range(25):
        side = rng.integers(-40, 41, 12)
        for i in range(12):
            a.coef_luma[c][i] = int(side[i])
        a.coef_luma[c][12] = int(512 - 2 * side.sum())
    side = rng.integers(-40, 41, 6)
    for i in range(6):
        a.coef_chroma[i] = int(side[i])
    a.coef_chroma[6] = int(512 - 2 * side.sum())
    for i in range(3):
        a.enable[i] = enable[i]
    return a
