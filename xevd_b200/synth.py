"""Synthetic pre-parsed CU arrays (BASELINE.json configs 2-5).

There are no EVC bitstreams on the build or GPU boxes, so workloads are generated the way SURVEY 8(d)
specifies: random reference pictures, a CU partition, quarter-pel motion vectors and coefficients that are
the quantised forward DCT of bounded residuals (so every inverse-transform sum stays in range, T4).

Deterministic: everything derives from the integer seed.
"""
from __future__ import annotations

import numpy as np

from .abi import CU_DTYPE, EXT_DTYPE, CUF_ATS_INTRA, CUF_CHROMA, CUF_LUMA, MODE_INTER, MODE_INTRA, make_params
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


def partition_btt(w: int, h: int, log2_ctu: int, rng, suco: bool = True, min_log2: int = 3, leaf_bias: float = 0.35):
    """Main-profile style partition: binary / ternary splits in both directions from the CTU (sps_btt_flag), aspect ratios up
    to 1:4, optional right-to-left child order for vertical splits (sps_suco_flag, src_main/xevdm.c:1854-1933), so that
    non-square CUs and CUs whose RIGHT neighbour is decoded first occur.  CUs with a dimension of 128 occur for 128x128 CTUs
    (they carry up to four 64x64 transform sub-blocks)."""
    out = []
    first = [0]
    ctu = 1 << log2_ctu

    def rec(x, y, lw, lh, depth):
        if x >= w or y >= h:
            return
        cw, ch = 1 << lw, 1 << lh
        over_x, over_y = x + cw > w, y + ch > h
        opts = []
        if lw - 1 >= min_log2 and lw - 1 >= lh - 2:
            opts.append("bv")
        if lh - 1 >= min_log2 and lh - 1 >= lw - 2:
            opts.append("bh")
        if lw - 2 >= min_log2 and lw - 2 >= lh - 2 and not over_x:
            opts.append("tv")
        if lh - 2 >= min_log2 and lh - 2 >= lw - 2 and not over_y:
            opts.append("th")
        must = over_x or over_y
        if must:
            opts = [o for o in opts if (o == "bv" and over_x) or (o == "bh" and over_y)] or opts
        if opts and (must or rng.random() >= min(1.0, leaf_bias * (depth + 1) * 0.5)):
            o = opts[int(rng.integers(0, len(opts)))]
            rtl = suco and o in ("bv", "tv") and rng.random() < 0.5
            if o == "bv":
                ch_ = [(x, y, lw - 1, lh), (x + (cw >> 1), y, lw - 1, lh)]
            elif o == "bh":
                ch_ = [(x, y, lw, lh - 1), (x, y + (ch >> 1), lw, lh - 1)]
            elif o == "tv":
                ch_ = [(x, y, lw - 2, lh), (x + (cw >> 2), y, lw - 1, lh), (x + 3 * (cw >> 2), y, lw - 2, lh)]
            else:
                ch_ = [(x, y, lw, lh - 2), (x, y + (ch >> 2), lw, lh - 1), (x, y + 3 * (ch >> 2), lw, lh - 2)]
            if rtl:
                ch_ = ch_[::-1]
            for c in ch_:
                rec(*c, depth + 1)
        else:
            out.append((x, y, lw, lh))

    for cy in range(0, h, ctu):
        for cx in range(0, w, ctu):
            rec(cx, cy, log2_ctu, log2_ctu, 0)
            first.append(len(out))
    return out, first


def make_inter_frame(w: int, h: int, *, bit_depth: int = 10, variant: str = "A", seed: int = 1,
                     n_refs: int = 1, bi_frac: float | None = None, coded_frac: float = 1.0,
                     mv_range_px: int = 32, resid_scale: float | None = None, iqt: bool = False,
                     log2_cu: int = 4, main_mv: bool = False, log2_ctu: int = 6, suco: bool = True, ats_inter_frac: float = 0.0,
                     min_log2: int = 3):
    """Config-2 style inter picture.

    variant "A": uniform 16x16 CUs, uni-prediction, every CU coded in all three planes.
    variant "B": random quadtree {64,32,16,8}, 50 % bi-prediction.
    variant "C": Main-profile binary/ternary partition (non-square CUs, optional SUCO order, CTU 32/64/128), 50 % bi-prediction;
                 CUs with a 128 dimension carry a random non-empty subset of their 64x64 transform sub-blocks.
    Returns (params, CuList).  Reference pictures are produced separately (make_refs).
    """
    rng = np.random.default_rng(seed)
    if variant == "A":
        parts, first = partition_uniform(w, h, log2_ctu, log2_cu)
        bi = 0.0 if bi_frac is None else bi_frac
    elif variant == "C":
        parts, first = partition_btt(w, h, log2_ctu, rng, suco=suco, min_log2=min_log2)
        bi = 0.5 if bi_frac is None else bi_frac
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
    lw_a, lh_a = cus["log2w"].astype(np.int64), cus["log2h"].astype(np.int64)
    if ats_inter_frac > 0:
        # ats_inter (sub-block transform, Main tool_ats): coded inter CUs up to 64x64; split kinds allowed by size as in
        # xevdm_check_ats_inter_info_coded (src_main/xevdm_util.c:3565-3583): halves need >= 8, quarters >= 16
        pick = rng.random(n) < ats_inter_frac
        for i in np.nonzero(pick & coded & (lw_a <= 6) & (lh_a <= 6))[0]:
            kinds = [k for k, ok in ((1, lw_a[i] >= 3), (2, lh_a[i] >= 3), (3, lw_a[i] >= 4), (4, lh_a[i] >= 4)) if ok]
            if kinds:
                cus["ats"][i] = (int(kinds[int(rng.integers(0, len(kinds)))]) << 2) | (int(rng.integers(0, 2)) << 5)
    # nnz_sub bits: one per 64x64 luma (32x32 chroma) transform sub-block, bit (j << 1) | i (xevd_eco.c:618-625)
    nsx, nsy = np.where(lw_a > 6, 2, 1), np.where(lh_a > 6, 2, 1)
    full = np.where(nsx == 2, np.where(nsy == 2, 0xF, 0x3), np.where(nsy == 2, 0x5, 0x1))
    sub = np.zeros((n, 3), np.int64)
    for c in range(3):
        pick = rng.integers(1, 16, n) & full
        sub[:, c] = np.where((full == 1) | (pick == 0), full, pick)
    cus["cbf"] = np.where(coded, sub[:, 0] | (sub[:, 1] << 4) | (sub[:, 2] << 8), 0).astype(np.uint16)

    # coefficient stream: quantised forward DCT of Laplacian residuals, per transform sub-block, stored CU-raster
    if resid_scale is None:
        # ~10 % non-zero levels: Laplacian scale ~ 0.22 of a quantiser step (dq_scale/64)
        resid_scale = 0.22 * dq_scale(qp_y, iqt) / 64.0
    from .frame import ats_inter_tu
    tu = np.array([ats_inter_tu(int(a_), int(w_), int(h_))[:2] for a_, w_, h_ in zip(cus["ats"], lw_a, lh_a)], np.int64).reshape(n, 2)
    tlw_a, tlh_a = tu[:, 0], tu[:, 1]                  # coded block = the TU (the whole CU without ats_inter)
    sizes = 1 << (tlw_a + tlh_a)
    a8 = lambda v: (v + 7) & ~7          # plane blocks padded to multiples of 8 int16
    per_cu = np.where(coded, a8(sizes) + 2 * a8(sizes // 4), 0)
    off = np.concatenate(([0], np.cumsum(per_cu)))
    cus["coef_off"] = off[:-1].astype(np.uint32)
    coef = np.zeros(int(off[-1]), np.int16)

    def plane_levels(m, pw, ph, bits, tmax):
        """levels of m planes of pw x ph, transform sub-blocks of at most tmax, uncoded sub-blocks zero"""
        tw, th = min(pw, tmax), min(ph, tmax)
        nx, ny = pw // tw, ph // th
        r = rng.laplace(0.0, resid_scale, (m, ny, nx, th, tw))
        lev = quantised_dct(r, qp_y, iqt)
        for j in range(ny):
            for i in range(nx):
                lev[:, j, i][((bits >> ((j << 1) | i)) & 1) == 0] = 0
        return lev.transpose(0, 1, 3, 2, 4).reshape(m, ph * pw)

    shapes = np.unique(np.stack([tlw_a, tlh_a], 1), axis=0)
    for lw_, lh_ in shapes:
        sel = np.nonzero((tlw_a == lw_) & (tlh_a == lh_) & coded)[0]
        if len(sel) == 0:
            continue
        cw_, ch_ = 1 << int(lw_), 1 << int(lh_)
        ly = plane_levels(len(sel), cw_, ch_, sub[sel, 0], 64)
        lu = plane_levels(len(sel), cw_ // 2, ch_ // 2, sub[sel, 1], 32)
        lv = plane_levels(len(sel), cw_ // 2, ch_ // 2, sub[sel, 2], 32)
        nc = a8(cw_ * ch_ // 4)
        lcp = np.zeros((len(sel), 2, nc), np.int16)
        lcp[:, 0, :lu.shape[1]] = lu
        lcp[:, 1, :lv.shape[1]] = lv
        lyp = np.zeros((len(sel), a8(cw_ * ch_)), np.int16)
        lyp[:, :ly.shape[1]] = ly
        blk = np.concatenate([lyp, lcp.reshape(len(sel), -1)], axis=1)
        idx = off[sel][:, None] + np.arange(blk.shape[1])[None, :]
        coef[idx] = blk
    prm = make_params(w, h, bit_depth=bit_depth, log2_ctu=log2_ctu, poc=8, tool_iqt=int(iqt), tool_admvp=int(main_mv),
                      tool_ats=int(ats_inter_frac > 0), tool_suco=int(variant == "C" and suco))
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
            pic.map_unrefined_mv[sl] = pic.map_mv[sl]
            pic.map_refi[sl] = -1 if intra else cu["refi"]
        qp = int(rng.integers(qp_lo, qp_hi + 1))
        m = (1 << 31) | (qp << 16) | (int(rng.random() < intra_frac) << 15) | ((int(cu["cbf"]) & 1) << 24)
        for j in range(nh):
            pic.map_scu[(y0 + j) * ws + x0:(y0 + j) * ws + x0 + nw] = m
    return pic


def split_local_dual_tree(cl: CuList, rng, frac: float = 0.5):
    """Local dual tree (Main, 4:2:0): where a split would leave chroma blocks narrower than 4 samples the decoder switches the
    node to intra-only, codes the split's leaves as luma-only CUs (TREE_L) and then the whole node once more as ONE chroma-only
    CU (TREE_C) (xevd_entropy_dec_tree / xevd_recon_tree, src_main/xevdm.c:1775-1846,1872-1927).  Turns a fraction of the CUs whose
    short side is 8 into such groups: luma leaves of 4 x N / N x 4 (halves; an 8x8 node may also quarter), flags = CUF_LUMA, then
    the node itself with flags = CUF_CHROMA.  Call BEFORE add_intra_cus (which makes every such CU intra and derives the
    chroma CU's DM mode from the luma leaf at the node's centre).  Coefficient blocks of a dual-tree CU hold its own planes only."""
    a8 = lambda v: (v + 7) & ~7
    old, coef_old = cl.cus, cl.coef
    rows, blocks, off = [], [], 0
    remap = {}

    def levels(n):
        v = rng.integers(-6, 7, n) * (rng.random(n) < 0.15)
        v[0] = int(rng.integers(-60, 61))
        out = np.zeros(a8(n), np.int16)
        out[:n] = v
        return out

    def emit(cu, block):
        nonlocal off
        cu = cu.copy()
        cu["coef_off"] = off
        rows.append(cu)
        blocks.append(block)
        off += len(block)

    for i in range(len(old)):
        remap[i] = len(rows)
        cu = old[i]
        lw, lh = int(cu["log2w"]), int(cu["log2h"])
        n = 1 << (lw + lh)
        coded = int(cu["cbf"]) != 0
        own = coef_old[int(cu["coef_off"]):int(old[i + 1]["coef_off"]) if i + 1 < len(old) else len(coef_old)]       # ats_inter CUs carry their TU only
        ats_inter = (int(cu["ats"]) >> 2) & 7
        if min(lw, lh) != 3 or max(lw, lh) > 4 or ats_inter or rng.random() >= frac:
            emit(cu, own)
            continue
        x, y = int(cu["x"]), int(cu["y"])
        if lw == 3 and lh == 3:
            kind = ("bv", "bh", "q")[int(rng.integers(0, 3))]
        else:
            kind = "bv" if lw == 3 else "bh"
        if kind == "bv":
            leaves = [(x, y, 2, lh), (x + 4, y, 2, lh)]
            if rng.random() < 0.5:
                leaves = leaves[::-1]                  # SUCO: right child first
        elif kind == "bh":
            leaves = [(x, y, lw, 2), (x, y + 4, lw, 2)]
        else:
            leaves = [(x, y, 2, 2), (x + 4, y, 2, 2), (x, y + 4, 2, 2), (x + 4, y + 4, 2, 2)]
        for (lx, ly, llw, llh) in leaves:
            c = cu.copy()
            c["x"], c["y"], c["log2w"], c["log2h"] = lx, ly, llw, llh
            c["flags"] = (int(cu["flags"]) & ~CUF_CHROMA) | CUF_LUMA
            c["ats"] = 0
            has = rng.random() < 0.7
            c["cbf"] = 1 if has else 0
            emit(c, levels(1 << (llw + llh)) if has else np.zeros(0, np.int16))
        c = cu.copy()
        c["flags"] = (int(cu["flags"]) & ~CUF_LUMA) | CUF_CHROMA
        c["ats"] = 0
        c["cbf"] = int(cu["cbf"]) & 0xff0
        emit(c, own[a8(n):] if coded else np.zeros(0, np.int16))
    remap[len(old)] = len(rows)
    cl.cus = np.array(rows, CU_DTYPE)
    cl.coef = np.concatenate(blocks).astype(np.int16) if blocks else np.zeros(0, np.int16)
    cl.ctu_first = np.array([remap[int(v)] for v in cl.ctu_first], np.uint32)
    return cl


def add_intra_cus(cl: CuList, rng, intra_frac: float = 1.0, n_modes: int = 5, constrained: bool = False, eipd: bool = False,
                  ats_intra_frac: float = 0.0, ibc_frac: float = 0.0):
    """Turn a fraction of the CUs of a picture into intra CUs (Baseline modes 0..4 for luma and chroma) and derive their
    neighbour-availability masks the way the decoder does: a neighbouring SCU is available when it lies inside the picture
    and was reconstructed earlier in decoding order (COD bit of map_scu; xevd_get_nbr_b, src_base/xevd_ipred.c:49-91).
    With `constrained` (pps.constrained_intra_pred_flag) it must also be intra.
    ibc_frac: fraction of the remaining inter CUs turned into intra-block-copy CUs (Main, sps ibc_flag) whose block vector points
    into the already decoded part of the current or the left neighbouring CTU (the legal range of EVC block vectors).
    eipd: Main-profile mode set (33 luma modes, chroma modes DM/BI/DC/HOR/VER = 0..4) plus the right-column mask and
    avail_lr (xevd_check_nev_avail, src_base/xevd_util.c:1156-1174) that xevdm_get_nbr / xevdm_ipred consume."""
    cus = cl.cus
    n = len(cus)
    w_scu, h_scu = (cl.w + 3) >> 2, (cl.h + 3) >> 2
    is_intra = (rng.random(n) < intra_frac) & (((cus["ats"] >> 2) & 7) == 0)       # ats_inter CUs stay inter
    dual = (cus["flags"] & (CUF_LUMA | CUF_CHROMA)) != (CUF_LUMA | CUF_CHROMA)       # local dual tree: intra only (split_local_dual_tree)
    is_intra |= dual
    if ats_intra_frac > 0:
        # ats_intra_cu (Main tool_ats): intra CUs of at most 32x32 with coded luma; ats mode = h << 1 | v (0 DST-7, 1 DCT-8)
        ok = is_intra & (cus["log2w"] <= 5) & (cus["log2h"] <= 5) & ((cus["cbf"] & 15) != 0) & (rng.random(n) < ats_intra_frac)
        cus["flags"] = np.where(ok, cus["flags"] | CUF_ATS_INTRA, cus["flags"])
        cus["ats"] = np.where(ok, rng.integers(0, 4, n), cus["ats"]).astype(np.uint8)
    cus["mode"] = np.where(is_intra, MODE_INTRA, MODE_INTER)
    ipm = rng.integers(0, n_modes, (n, 2)).astype(np.int8)
    if eipd:
        ipm[:, 0] = rng.integers(0, 33, n)
        ipm[:, 1] = rng.integers(0, 5, n)
    cus["refi"] = np.where(is_intra[:, None], ipm, cus["refi"])
    cus["mv"][is_intra] = 0
    ext = [np.zeros(1, EXT_DTYPE)[0]]
    cod = np.zeros((h_scu, w_scu), bool)
    intra_map = np.zeros((h_scu, w_scu), bool)
    ipm_map = np.zeros((h_scu, w_scu), np.int8)
    from .abi import MODE_IBC
    ctu_scu = (1 << cl.log2_ctu) >> 2
    want_ibc = (~is_intra) & (((cus["ats"] >> 2) & 7) == 0) & (rng.random(n) < ibc_frac) if ibc_frac > 0 else np.zeros(n, bool)
    if ibc_frac > 0 and dual.any():
        # luma-only leaves of a dual tree node may be IBC CUs (intra-only nodes allow it; xevdm_IBC_mc copies luma only for TREE_L,
        # src_main/xevdm_mc.c:2059-2073); drawn separately so that pictures without dual tree nodes stay what they were
        want_ibc |= ((cus["flags"] & (CUF_LUMA | CUF_CHROMA)) == CUF_LUMA) & (rng.random(n) < ibc_frac)
    for i in range(n):
        cu = cus[i]
        xs, ys = int(cu["x"]) >> 2, int(cu["y"]) >> 2
        nw, nh = 1 << (int(cu["log2w"]) - 2), 1 << (int(cu["log2h"]) - 2)
        if want_ibc[i]:
            # candidate source blocks: SCU-aligned or not, fully decoded, inside the current / left CTU of this CTU row
            x0, y0, cw_, ch_ = int(cu["x"]), int(cu["y"]), 4 * nw, 4 * nh
            cy0 = (ys // ctu_scu) * ctu_scu * 4
            cx0 = max(0, ((xs // ctu_scu) - 1) * ctu_scu * 4)
            for _ in range(24):
                sx = int(rng.integers(cx0, x0 + cw_))
                sy = int(rng.integers(cy0, min(cy0 + 4 * ctu_scu, cl.h) - ch_ + 1))
                if sx + cw_ > cl.w or sx < 1 or sy < 1:
                    continue
                # odd vectors floor towards the upper left for chroma (>> 1): require one more decoded sample there
                if cod[(sy - 1) >> 2:((sy + ch_ - 1) >> 2) + 1, (sx - 1) >> 2:((sx + cw_ - 1) >> 2) + 1].all():
                    cu["mode"] = MODE_IBC
                    cu["refi"] = -1
                    cu["mv"] = 0
                    cu["mv"][0] = (sx - x0, sy - y0)
                    cu["flags"] = int(cu["flags"]) & ~CUF_ATS_INTRA
                    cu["ats"] = 0
                    is_intra[i] = False
                    break
        if is_intra[i] and not (int(cu["flags"]) & CUF_LUMA):
            # TREE_C CU: ipm[0] comes from map_ipm at the node's centre SCU, IPD_DC if that SCU is not intra (src_main/xevdm.c:1081-1092)
            yc, xc = ys + (nh >> 1), xs + (nw >> 1)
            cu["refi"][0] = ipm_map[yc, xc] if intra_map[yc, xc] else 0
        if is_intra[i]:
            ok = (lambda yy, xx: cod[yy, xx] and (not constrained or intra_map[yy, xx]))
            up = left = right = 0
            for k in range(nw + nh):
                if ys > 0 and xs + k < w_scu and ok(ys - 1, xs + k):
                    up |= 1 << k
                if xs > 0 and ys + k < h_scu and ok(ys + k, xs - 1):
                    left |= 1 << k
                if xs + nw < w_scu and ys + k < h_scu and ok(ys + k, xs + nw):
                    right |= 1 << k
            ul = int(ys > 0 and xs > 0 and ok(ys - 1, xs - 1))
            # avail_lr looks at COD only (no constrained-intra test)
            lr = int(xs > 0 and cod[ys, xs - 1]) | (int(xs + nw < w_scu and cod[ys, xs + nw]) << 1)
            e = np.zeros(1, EXT_DTYPE)[0]
            e["q"][0], e["q"][1], e["q"][2] = up, left, right
            cu["avail"] = lr | (ul << 2)
            cu["mv"][1] = np.frombuffer(np.uint32(len(ext)).tobytes(), np.int16)
            ext.append(e)
        cod[ys:ys + nh, xs:xs + nw] = True
        if int(cu["flags"]) & CUF_LUMA:               # xevdm_set_dec_info publishes for luma-carrying CUs only (xevdm_util.c:4241)
            intra_map[ys:ys + nh, xs:xs + nw] = is_intra[i]
            ipm_map[ys:ys + nh, xs:xs + nw] = int(cu["refi"][0]) if is_intra[i] else 0
    cl.ext = np.array(ext, EXT_DTYPE)
    return cl


def add_affine_cus(cl: CuList, rng, frac: float = 0.5):
    """Turn a fraction of the inter CUs of at least 8x8 into affine CUs (Main tool_affine): control point vectors = the CU's vectors
    plus per-vertex offsets drawn from a wide range, so that the 8x8-or-larger sub-block path, the per-sample EIF path (with and
    without the clamped vector window) and 4- / 6-parameter models all occur; extension records carry the control points and the
    value xevdm_set_dec_info would publish as unrefined vector."""
    from .abi import CUF_AFF6, MODE_AFFINE
    cus = cl.cus
    ext = list(cl.ext)
    scales = np.array([0, 0, 1, 2, 4, 10, 40])
    for i in range(len(cus)):
        cu = cus[i]
        if int(cu["mode"]) != MODE_INTER or int(cu["log2w"]) < 3 or int(cu["log2h"]) < 3 or rng.random() >= frac:
            continue
        e = np.zeros(1, EXT_DTYPE)[0]
        rec = np.zeros(16, np.int16)            # cp[2][3][2] then mv_unref[2][2]
        sc = int(scales[int(rng.integers(0, len(scales)))])
        for l in range(2):
            base = cu["mv"][l].astype(np.int64)
            for v in range(3):
                d = rng.integers(-sc, sc + 1, 2) if (v and sc) else np.zeros(2, np.int64)
                rec[(l * 3 + v) * 2:(l * 3 + v) * 2 + 2] = base + d
        rec[12:16] = rng.integers(-64, 65, 4)
        six = rng.random() < 0.5
        if six and rng.random() < 0.3:
            # strong horizontal stretch with a flat vertical gradient: EIF stays applicable but its fetch area exceeds the
            # memory-bandwidth bound, which switches on the clamped vector window (eif_derive_mv_clip_range)
            for l in range(2):
                rec[(l * 3 + 1) * 2] += (3 + int(rng.integers(0, 3))) << int(cu["log2w"])
                rec[(l * 3 + 2) * 2:(l * 3 + 2) * 2 + 2] = rec[(l * 3) * 2:(l * 3) * 2 + 2]
        e["q"] = rec.view(np.uint64)
        cu["mode"] = MODE_AFFINE
        if six:
            cu["flags"] = int(cu["flags"]) | CUF_AFF6
        cu["mv"] = 0
        cu["mv"][1] = np.frombuffer(np.uint32(len(ext)).tobytes(), np.int16)
        ext.append(e)
    cl.ext = np.array(ext, EXT_DTYPE)
    return cl


def derive_avail_cu(cl: CuList):
    """XB200_CU.avail_cu for every CU: what xevd_get_avail_intra (src_base/xevd_util.c:689-745) returns when the CU is reached in
    decoding order (one tile, one slice) - consumed by the HTDF ring fetch (src_main/xevdm_recon.c:299-385)"""
    w_scu, h_scu = (cl.w + 3) >> 2, (cl.h + 3) >> 2
    cod = np.zeros((h_scu, w_scu), bool)
    for cu in cl.cus:
        xs, ys = int(cu["x"]) >> 2, int(cu["y"]) >> 2
        nw, nh = 1 << (int(cu["log2w"]) - 2), 1 << (int(cu["log2h"]) - 2)
        a = 0
        if xs > 0 and cod[ys, xs - 1]:
            a |= 1 << 1                                                           # AVAIL_LE
            if ys + nh + nw - 1 < h_scu and cod[ys + nh + nw - 1, xs - 1]:
                a |= 1 << 7                                                       # AVAIL_LO_LE
        if ys > 0:
            a |= (1 << 0) | (1 << 9)                                              # AVAIL_UP, AVAIL_RI_UP (single tile)
            if xs > 0 and cod[ys - 1, xs - 1]:
                a |= 1 << 5                                                       # AVAIL_UP_LE
            if xs + nw < w_scu and cod[ys - 1, xs + nw]:
                a |= 1 << 6                                                       # AVAIL_UP_RI
        if xs + nw < w_scu and cod[ys, xs + nw]:
            a |= 1 << 3                                                           # AVAIL_RI
            if ys + nh + nw - 1 < h_scu and cod[ys + nh + nw - 1, xs + nw]:
                a |= 1 << 8                                                       # AVAIL_LO_RI
        cu["avail_cu"] = a
        cod[ys:ys + nh, xs:xs + nw] = True
    return cl


def make_dmvr_case(w: int, h: int, *, bit_depth: int = 10, variant: str = "C", seed: int = 1, flag_frac: float = 0.8, noise: int = 3, **kw):
    """Picture exercising decoder-side motion vector refinement (Main tool_dmvr): two reference pictures at equal POC distance
    on either side of the current one whose content is the same texture displaced by a few samples (plus noise), bi-predicted
    CUs with roughly mirrored vectors that are off by up to two samples, so the SAD search moves, stops early on (near-)zero
    cost and takes the parabolic sub-sample step; XB200_CUF_DMVR is set on a fraction of the CUs (also on some that fail
    xevdm_mc's own conditions: uni-prediction, both lists on the same side, CUs narrower than 8)."""
    from .abi import CUF_DMVR
    rng = np.random.default_rng(seed + 7000)
    prm, cl = make_inter_frame(w, h, bit_depth=bit_depth, variant=variant, seed=seed, n_refs=2, bi_frac=0.8, mv_range_px=6, **kw)
    prm.tool_dmvr = 1
    cus = cl.cus
    n = len(cus)
    bi = (cus["refi"][:, 0] >= 0) & (cus["refi"][:, 1] >= 0)
    # mirrored motion with a small error for most bi-predicted CUs
    mirror = bi & (rng.random(n) < 0.85)
    err = rng.integers(-8, 9, (n, 2)).astype(np.int16)
    cus["mv"][mirror, 1, :] = -cus["mv"][mirror, 0, :] + err[mirror]
    cus["flags"] = np.where(rng.random(n) < flag_frac, cus["flags"] | CUF_DMVR, cus["flags"]).astype(np.uint8)
    # reference pictures: one smooth random texture per plane; picture 1 shows it displaced by (3, -2) luma samples, plus noise
    hi = (1 << bit_depth) - 1

    def texture(hh, ww):
        base = rng.integers(0, hi + 1, (hh // 4 + 6, ww // 4 + 6)).astype(np.float64)
        up = np.kron(base, np.ones((4, 4)))
        k = np.ones(5) / 5.0
        up = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, up)
        return np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, up)

    tex = [texture(h, w), texture(h // 2, w // 2), texture(h // 2, w // 2)]
    pics = []
    for poc, (sx, sy) in ((0, (0, 0)), (16, (3, -2))):
        p = HostPicture(w, h, poc)
        for k_, pl in enumerate(p.planes()):
            dx_, dy_ = (sx, sy) if k_ == 0 else (sx // 2, sy // 2)
            v = tex[k_][8 + dy_:8 + dy_ + pl.shape[0], 8 + dx_:8 + dx_ + pl.shape[1]]
            if poc:
                v = v + rng.integers(-noise, noise + 1, pl.shape)
            pl[...] = np.clip(np.rint(v), 0, hi).astype(np.int16)
        pics.append(p.pad_borders())
    return prm, cl, pics


def make_main_frame(w: int, h: int, *, bit_depth: int = 10, seed: int = 1, log2_ctu: int = 6, slice_qp: int = 34, intra_frac: float = 0.25,
                    iqt: bool = True, coded_frac: float = 0.6):
    """BASELINE config 3 in miniature: one Main-profile picture with every hot-path tool on - binary/ternary partition with SUCO order,
    1/16-pel interpolation, IQT, ATS (intra + sub-block inter), EIPD intra, IBC, HTDF, DMVR, affine - plus the parameters of the
    picture-wide passes (ADDB deblocking, ALF).  Returns (params, CuList, reference pictures, ALF parameters, ALF CTB flags)."""
    rng = np.random.default_rng(seed + 9000)
    prm, cl, refs = make_dmvr_case(w, h, bit_depth=bit_depth, variant="C", seed=seed, noise=3, coded_frac=coded_frac, main_mv=True,
                                   ats_inter_frac=0.3, iqt=iqt, log2_ctu=log2_ctu)
    prm.tool_eipd = prm.tool_ibc = prm.tool_htdf = prm.tool_affine = prm.tool_addb = prm.tool_alf = 1
    prm.slice_qp = slice_qp
    prm.qp_u_offset, prm.qp_v_offset = 1, -2
    cl.cus["qp_map"] = rng.integers(24, 46, cl.n_cu)
    add_intra_cus(cl, rng, intra_frac, eipd=True, ats_intra_frac=0.5, ibc_frac=0.15)
    add_affine_cus(cl, rng, 0.3)
    derive_avail_cu(cl)
    ctu = 1 << log2_ctu
    n_ctu = ((w + ctu - 1) // ctu) * ((h + ctu - 1) // ctu)
    return prm, cl, refs, make_alf_params(rng), (rng.random(n_ctu) < 0.8).astype(np.uint8)


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


def make_dra_params(rng, bit_depth: int = 10):
    """random but plausible DRA look-up tables (what xevd_init_dra builds from an APS): a monotonic luma mapping within the 10-bit
    range and chroma scale factors around 1.0 in 9 fractional bits"""
    from .abi import DraParams
    d = DraParams()
    steps = rng.integers(0, 3, 1024)
    lut = np.minimum(np.cumsum(steps) * 1023 // max(1, int(steps.sum())), 1023)
    for i in range(1024):
        d.luma_inv_scale_lut[i] = int(lut[i])
    for c in range(2):
        sc = np.clip(512 + np.cumsum(rng.integers(-3, 4, 1024)), 256, 1023)
        for i in range(1024):
            d.chroma_inv_scale_lut[c][i] = int(sc[i])
    return d
