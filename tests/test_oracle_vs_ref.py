"""Pins the CPU oracle (oracle/orc_*.c) against the UNMODIFIED reference compiled from /root/reference
(oracle/_ref/libxevd_ref.so): tables, leaf kernels and CU-level picture reconstruction.  C, SSE and AVX2
variants of the reference are all exercised."""
import numpy as np
import pytest

from xevd_b200 import synth
from xevd_b200.frame import HostPicture


def test_dct2_tables(oracle, reference):
    for lg in range(1, 7):
        assert np.array_equal(oracle.dct2_matrix(lg), reference.dct2_matrix(lg)), f"tm{1 << lg}"


def test_ats_tables(oracle, reference):
    for lg in range(2, 6):
        for dst7 in (0, 1):
            assert np.array_equal(oracle.ats_matrix(dst7, lg), reference.ats_matrix(dst7, lg))


def test_mc_taps(oracle, reference):
    for m in (0, 1):
        lo, co = oracle.mc_taps(m)
        lr, cr = reference.mc_taps(m)
        assert np.array_equal(lo, lr) and np.array_equal(co, cr)


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_mc_leaf(oracle, reference, impl, bit_depth):
    reference.set_impl(impl)
    rng = np.random.default_rng(100 + impl + bit_depth)
    plane = rng.integers(0, 1 << bit_depth, (200, 232), dtype=np.int16)
    org = (40, 36)
    sizes = [4, 8, 16, 32, 64, 128]
    for t in range(250):
        main = bool(t & 1)
        w, h = sizes[rng.integers(6)], sizes[rng.integers(6)]
        w, h = min(w, 128), min(h, 96)
        # luma: 1/16 pel; Baseline only has quarter-pel phases
        step = 1 if main else 4
        fx, fy = int(rng.integers(0, 16 // step)) * step, int(rng.integers(0, 16 // step)) * step
        gx, gy = (int(rng.integers(-20, 40)) << 4) + fx, (int(rng.integers(-20, 40)) << 4) + fy
        # the variant comes from the unclipped vector; make it sometimes disagree with the phase (T3)
        ox, oy = (fx, fy) if rng.random() < 0.8 else (int(rng.integers(0, 16)), int(rng.integers(0, 16)))
        a = oracle.mc(plane, org, (gx, gy), (ox, oy), w, h, bit_depth, False, main)
        b = reference.mc(plane, org, (gx, gy), (ox, oy), w, h, bit_depth, False, main)
        assert np.array_equal(a, b), (impl, "luma", w, h, gx, gy, ox, oy, main)
        cw, ch = max(2, w // 2), max(2, h // 2)
        cstep = 1 if main else 4
        cfx, cfy = int(rng.integers(0, 32 // cstep)) * cstep, int(rng.integers(0, 32 // cstep)) * cstep
        cgx, cgy = (int(rng.integers(-10, 30)) << 5) + cfx, (int(rng.integers(-10, 30)) << 5) + cfy
        cox, coy = (cfx, cfy) if rng.random() < 0.8 else (int(rng.integers(0, 32)), int(rng.integers(0, 32)))
        a = oracle.mc(plane, org, (cgx, cgy), (cox, coy), cw, ch, bit_depth, True, main)
        b = reference.mc(plane, org, (cgx, cgy), (cox, coy), cw, ch, bit_depth, True, main)
        assert np.array_equal(a, b), (impl, "chroma", cw, ch, cgx, cgy, cox, coy, main)
    reference.set_impl(2)


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("iqt", [0, 1])
def test_itdq_all_shapes(oracle, reference, impl, iqt):
    reference.set_impl(impl)
    rng = np.random.default_rng(7 + impl)
    for bit_depth in (8, 10):
        qp = 32 + 6 * (bit_depth - 8)
        for lw in range(1, 7):
            for lh in range(1, 7):
                w, h = 1 << lw, 1 << lh
                for rep in range(3):
                    res = rng.laplace(0, 40.0 * (1 << (bit_depth - 8)), (h, w))
                    lev = synth.quantised_dct(res, qp, bool(iqt))
                    if rep == 2:      # sparse high-frequency content too
                        lev[:] = 0
                        # (IQT: stay inside the 32 lowest frequencies, see synth.quantised_dct)
                        lim = 32 if iqt else 64
                        lev[rng.integers(min(h, lim)), rng.integers(min(w, lim))] = rng.integers(-40, 40)
                    a = oracle.itdq_block(lev, qp, bit_depth, iqt)
                    b = reference.itdq_block(lev, qp, bit_depth, iqt)
                    assert np.array_equal(a, b), (impl, iqt, bit_depth, lw, lh, rep)
    reference.set_impl(2)


@pytest.mark.parametrize("variant,bit_depth,iqt", [("A", 10, 0), ("B", 10, 0), ("B", 8, 0), ("B", 10, 1)])
def test_recon_frame_inter(oracle, reference, variant, bit_depth, iqt):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bit_depth, variant=variant, seed=5, n_refs=2,
                                     coded_frac=0.8, mv_range_px=200, iqt=bool(iqt))
    cl.validate()
    refs = synth.make_refs(w, h, bit_depth, 2, seed=9)
    refs[1].poc = refs[0].poc if variant == "B" else refs[1].poc   # exercise the identical-motion shortcut
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), name


def test_pad(oracle, reference):
    rng = np.random.default_rng(3)
    p = HostPicture.random(72, 40, 10, rng)
    a = oracle.pad(p.copy())
    b = reference.pad(p.copy())
    assert np.array_equal(a.buf_y, b.buf_y) and np.array_equal(a.buf_u, b.buf_u) and np.array_equal(a.buf_v, b.buf_v)
    assert np.array_equal(a.buf_y, p.copy().pad_borders().buf_y)


@pytest.mark.parametrize("variant,log2_cu,bd,main_tbl", [("B", 4, 10, 0), ("A", 2, 10, 0), ("A", 3, 8, 1), ("B", 4, 8, 0), ("A", 6, 10, 1)])
def test_deblock_baseline_filter(oracle, reference, variant, log2_cu, bd, main_tbl):
    """Baseline deblocking filter (tool_addb = 0) through the reference's CU walkers vs the oracle, all four strength
    classes, QP 18..51, chroma QP offsets, 4x4 CUs (order-dependent chroma chains) and 64x64 CUs"""
    w, h = 192, 136
    rng = np.random.default_rng(40 + log2_cu + bd)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=31, n_refs=2, coded_frac=0.5, log2_cu=log2_cu,
                                     bi_frac=0.3, mv_range_px=3)
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    refs = synth.make_refs(w, h, bd, 2, seed=32)
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    # smooth the picture so that edge steps are in the range the filter acts on
    for pl in base.planes():
        pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
    synth.randomize_deblock_maps(base, cl, rng)
    tbl = synth.chroma_qp_table(bool(main_tbl))
    a = oracle.deblock_frame(prm, base.copy(), cl, tbl)
    b = reference.deblock_frame(prm, base.copy(), cl, tbl)
    changed = sum(int((x != y).sum()) for x, y in zip(a.planes(), base.planes()))
    assert changed > 500, "test picture does not exercise the filter"
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("variant,log2_cu,bd,intra_frac", [("B", 4, 10, 1.0), ("A", 2, 10, 1.0), ("B", 4, 8, 0.4), ("A", 6, 10, 1.0), ("A", 3, 10, 0.5)])
def test_recon_frame_intra_baseline(oracle, reference, variant, log2_cu, bd, intra_frac):
    """Baseline intra CUs (5 modes, luma + chroma) inside I and mixed pictures.  The reference derives neighbour availability
    from its own COD bits in decoding order; the oracle consumes the masks synth.add_intra_cus derived - so this also pins the
    mask derivation."""
    w, h = 200, 136
    rng = np.random.default_rng(7)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=81, n_refs=2, coded_frac=0.7, log2_cu=log2_cu)
    synth.add_intra_cus(cl, rng, intra_frac)
    refs = synth.make_refs(w, h, bd, 2, seed=82)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("variant,log2_cu,bd,aoff,boff", [("B", 4, 10, 0, 0), ("A", 2, 10, 2, -2), ("A", 3, 8, -3, 4), ("B", 4, 8, 6, 6), ("A", 6, 10, 0, 0), ("B", 4, 12, 0, 2)])
def test_deblock_addb(oracle, reference, variant, log2_cu, bd, aoff, boff):
    """Main-profile deblocking (tool_addb): bS 0..4 incl. the cross-CTU intra case, picture (not index) comparison with two
    reference indices aliasing one picture, alpha/beta offsets (negative ones wrap in the reference's u8 arguments)"""
    w, h = 192, 136
    rng = np.random.default_rng(140 + log2_cu + bd)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=33, n_refs=3, coded_frac=0.4, log2_cu=log2_cu,
                                     bi_frac=0.4, mv_range_px=2)
    prm.tool_addb = 1
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    prm.deblock_alpha_offset, prm.deblock_beta_offset = aoff, boff
    refs = synth.make_refs(w, h, bd, 3, seed=34)
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl in base.planes():
        pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
    synth.randomize_deblock_maps(base, cl, rng, intra_frac=0.15)
    tbl = synth.chroma_qp_table(True)
    ids = ((0, 1, 0), (2, 1, 0))          # L0 index 2 aliases picture 0; L1 is a permutation
    a = oracle.deblock_frame(prm, base.copy(), cl, tbl, True, ids)
    b = reference.deblock_frame(prm, base.copy(), cl, tbl, True, ids)
    changed = sum(int((x != y).sum()) for x, y in zip(a.planes(), base.planes()))
    assert changed > 300, "test picture does not exercise the filter"
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("w,h,bd,log2_ctu,enable", [(192, 136, 10, 6, (1, 1, 1)), (200, 72, 8, 6, (1, 0, 1)), (128, 128, 10, 5, (1, 1, 0)), (72, 200, 10, 6, (0, 1, 1)), (256, 136, 10, 7, (1, 1, 1))])
def test_alf(oracle, reference, w, h, bd, log2_ctu, enable):
    """adaptive loop filter: classification, 7x7 / 5x5 diamonds, mirrored margins at picture edges, partial CTUs, per-CTU flags"""
    rng = np.random.default_rng(w + h + bd)
    p = HostPicture.random(w, h, bd, rng)
    # structured content so that all direction classes occur: gradients + texture
    yy, xx = np.mgrid[0:h, 0:w]
    p.y[...] = np.clip((p.y.astype(np.int32) >> 3) + ((xx * 3 + yy * 5) % 97) * (1 << (bd - 8)) + ((xx // 8 + yy // 8) % 2) * (40 << (bd - 8)), 0, (1 << bd) - 1).astype(np.int16)
    prm = __import__("xevd_b200.abi", fromlist=["make_params"]).make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, enable)
    n_ctu = ((w + (1 << log2_ctu) - 1) >> log2_ctu) * ((h + (1 << log2_ctu) - 1) >> log2_ctu)
    flags = (rng.random(n_ctu) < 0.8).astype(np.uint8)
    a = oracle.alf_frame(prm, p.copy(), alf, flags)
    b = reference.alf_frame(prm, p.copy(), alf, flags)
    assert sum(int((x != y).sum()) for x, y in zip(a.planes(), p.planes())) > 1000
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


MAIN_PART_CASES = [("C", {}, 10), ("C", dict(log2_ctu=7), 10), ("C", dict(log2_ctu=5), 8), ("C", dict(suco=False), 10)]


@pytest.mark.parametrize("variant,kw,bd", MAIN_PART_CASES)
def test_recon_frame_inter_btt(oracle, reference, variant, kw, bd):
    """Main-profile partitions: non-square CUs (1:2, 1:4), 32/64/128 CTUs, CUs of 128 with 64x64 transform sub-blocks gated by
    nnz_sub (xevdm_sub_block_itdq, src_main/xevdm_itdq.c:790-887)"""
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=31, n_refs=2, coded_frac=0.8, **kw)
    cl.validate()
    refs = synth.make_refs(w, h, bd, 2, seed=32)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("variant,kw,bd,intra_frac", [("C", {}, 10, 1.0), ("C", dict(log2_ctu=7), 10, 1.0), ("C", dict(log2_ctu=5), 8, 0.5),
                                                      ("B", {}, 10, 1.0), ("A", dict(log2_cu=2), 10, 1.0), ("C", dict(suco=False), 12, 0.7)])
def test_recon_frame_intra_eipd(oracle, reference, variant, kw, bd, intra_frac):
    """Main-profile intra (tool_eipd): xevdm_get_nbr + xevdm_ipred / xevdm_ipred_uv, 33 luma and 5 chroma modes, all four avail_lr
    cases (right neighbours decoded first under SUCO order).  The reference derives availability and avail_lr from its own COD
    bits; the oracle consumes the masks of synth.add_intra_cus."""
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=21, n_refs=2, coded_frac=0.7, **kw)
    prm.tool_eipd = 1
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True)
    if variant == "C" and kw.get("suco", True):
        lr = set((cl.cus["avail"][cl.cus["mode"] == 0] & 3).tolist())
        assert lr == {0, 1, 2, 3}, "test picture does not reach every avail_lr case"
    refs = synth.make_refs(w, h, bd, 2, seed=9)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


ATS_CASES = [("C", {}, 10, 0.5, 0), ("C", dict(log2_ctu=7), 10, 0.3, 1), ("C", dict(log2_ctu=5), 8, 0.5, 0), ("B", {}, 10, 0.0, 1), ("A", dict(log2_cu=3), 10, 1.0, 0)]


@pytest.mark.parametrize("variant,kw,bd,intra_frac,iqt", ATS_CASES)
def test_recon_frame_ats(oracle, reference, variant, kw, bd, intra_frac, iqt):
    """Main tool_ats: ats_intra (DST-7 / DCT-8 luma transforms of intra CUs, xevdm_it_MxN_ats_intra) and ats_inter (sub-block
    transform: half / quarter TU at either side, position-dependent kernels, TU placement by xevdm_recon), combined with both
    DCT-2 flavours (Baseline and IQT) for the remaining blocks"""
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=41, n_refs=2, coded_frac=0.8, ats_inter_frac=0.6, iqt=bool(iqt), **kw)
    prm.tool_eipd = 1
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True, ats_intra_frac=0.7)
    cl.validate()
    assert (((cl.cus["ats"] >> 2) & 7) != 0).sum() > 10
    refs = synth.make_refs(w, h, bd, 2, seed=9)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


# min_log2=2: 4-wide / 4-high CUs in SUCO order - chroma edges 2 samples apart whose filtering order is the CUs' decoding order
DBK_MAIN_CASES = [({}, 10, 1), (dict(log2_ctu=7), 10, 1), (dict(log2_ctu=5), 8, 1), ({}, 10, 0), (dict(log2_ctu=7), 8, 0),
                  (dict(min_log2=2), 10, 0), (dict(min_log2=2, log2_ctu=5), 8, 0), (dict(min_log2=2, log2_ctu=7), 10, 0), (dict(min_log2=2), 8, 1)]


def deblock_main_inputs(oracle, kw, bd, addb):
    w, h = 256, 136
    rng = np.random.default_rng(170 + bd + addb)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=35, n_refs=3, coded_frac=0.4, bi_frac=0.4, mv_range_px=2,
                                     ats_inter_frac=0.3, **kw)
    prm.tool_addb = addb
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    refs = synth.make_refs(w, h, bd, 3, seed=34)
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl in base.planes():
        pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
    synth.randomize_deblock_maps(base, cl, rng, intra_frac=0.15)
    return w, h, prm, cl, base, synth.chroma_qp_table(True), ((0, 1, 0), (2, 1, 0))


@pytest.mark.parametrize("kw,bd,addb", DBK_MAIN_CASES)
def test_deblock_main_partitions(oracle, reference, kw, bd, addb):
    """both deblocking filters on Main-profile partitions: non-square CUs, edges off the 8x8 grid (ternary splits), 128-sample CUs
    with their 64-sample transform edge, and ats_inter CUs (ats_present raises the ADDB strength to 'coded', xevdm_df.c:415,902)"""
    w, h, prm, cl, base, tbl, ids = deblock_main_inputs(oracle, kw, bd, addb)
    a = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    b = reference.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    changed = sum(int((x != y).sum()) for x, y in zip(a.planes(), base.planes()))
    assert changed > 300, "test picture does not exercise the filter"
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


def tile_grid(w, h, log2_ctu, n_cols, n_rows):
    """uniform tile spacing as set_tile_info computes it (src_main/xevdm.c:2247-2258)"""
    wc, hc = (w + (1 << log2_ctu) - 1) >> log2_ctu, (h + (1 << log2_ctu) - 1) >> log2_ctu
    return [i * wc // n_cols for i in range(n_cols + 1)], [j * hc // n_rows for j in range(n_rows + 1)]


@pytest.mark.parametrize("across", [0, 1])
@pytest.mark.parametrize("kw,bd,addb,grid", [({}, 10, 1, (2, 2)), (dict(log2_ctu=5), 8, 1, (3, 2)), ({}, 10, 0, (2, 2)), (dict(log2_ctu=5, min_log2=2), 10, 0, (4, 3)),
                                             (dict(log2_ctu=7), 8, 0, (2, 1))])
def test_deblock_tiles(oracle, reference, kw, bd, addb, grid, across):
    """pictures of several tiles: an edge between two tiles is filtered only with loop_filter_across_tiles_enabled_flag
    (map_tidx tests of xevdm_df.c:142,233,274,877,1088,1106), both filters, SUCO partitions down to 4-wide CUs"""
    w, h, prm, cl, base, tbl, ids = deblock_main_inputs(oracle, kw, bd, addb)
    cb, rb = tile_grid(w, h, prm.log2_ctu, *grid)
    one = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    for o in (oracle, reference):
        o.set_tiles(cb, rb, bool(across))
    try:
        a = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
        b = reference.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    finally:
        for o in (oracle, reference):
            o.set_tiles()
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    differs = sum(int((x != y).sum()) for x, y in zip(a.planes(), one.planes()))
    assert (differs == 0) if across else (differs > 50), "the tile boundaries of the test picture carry no filtered edge"


@pytest.mark.parametrize("across", [0, 1])
@pytest.mark.parametrize("w,h,bd,log2_ctu,grid", [(256, 136, 10, 6, (2, 2)), (200, 120, 8, 5, (3, 2)), (384, 256, 10, 7, (3, 1)), (320, 192, 10, 5, (4, 4))])
def test_alf_tiles(oracle, reference, w, h, bd, log2_ctu, grid, across):
    """ALF on pictures of several tiles (alf_process_tile per tile): windows from the tile's own extended copy, margins mirrored at tile
    borders without the flag, replicated with it (and then mirrored only at the picture's left and top)"""
    rng = np.random.default_rng(w + h + bd + across)
    p = HostPicture.random(w, h, bd, rng)
    prm = __import__("xevd_b200.abi", fromlist=["make_params"]).make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, (1, 1, 1))
    n_ctu = ((w + (1 << log2_ctu) - 1) >> log2_ctu) * ((h + (1 << log2_ctu) - 1) >> log2_ctu)
    flags = (rng.random(n_ctu) < 0.8).astype(np.uint8)
    cb, rb = tile_grid(w, h, log2_ctu, *grid)
    one = oracle.alf_frame(prm, p.copy(), alf, flags)
    for o in (oracle, reference):
        o.set_tiles(cb, rb, bool(across))
    try:
        a = oracle.alf_frame(prm, p.copy(), alf, flags)
        b = reference.alf_frame(prm, p.copy(), alf, flags)
    finally:
        for o in (oracle, reference):
            o.set_tiles()
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()), np.argwhere(pa != pb)[:4].tolist())
    assert sum(int((x != y).sum()) for x, y in zip(a.planes(), one.planes())) > 100, "tile borders change nothing in this picture"


def dual_tree_inputs(variant, kw, bd, eipd, htdf, intra_frac=1.0, ibc=0.0):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=81, n_refs=2, coded_frac=0.7, **kw)
    prm.tool_eipd, prm.tool_htdf, prm.slice_qp, prm.tool_ibc = eipd, htdf, 37, int(ibc > 0)
    synth.split_local_dual_tree(cl, np.random.default_rng(5), 0.6)
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=bool(eipd), ibc_frac=ibc)
    synth.derive_avail_cu(cl)
    cl.validate()
    fl = cl.cus["flags"] & 3
    assert (fl == 1).sum() > 8 and (fl == 2).sum() > 3, "test picture has no dual-tree groups"
    return w, h, prm, cl, synth.make_refs(w, h, bd, 2, seed=9)


DUAL_TREE_CASES = [("C", {}, 10, 0, 0, 1.0), ("C", {}, 10, 1, 0, 1.0), ("C", dict(log2_ctu=5), 8, 1, 1, 1.0), ("A", dict(log2_cu=3), 10, 1, 1, 1.0),
                   ("C", dict(suco=False), 12, 0, 0, 0.5), ("C", dict(log2_ctu=7), 10, 1, 1, 0.4)]


@pytest.mark.parametrize("variant,kw,bd,eipd,htdf,intra_frac", DUAL_TREE_CASES)
def test_recon_frame_dual_tree(oracle, reference, variant, kw, bd, eipd, htdf, intra_frac):
    """local dual tree (src_main/xevdm.c:1828-1846,1908-1927): luma-only leaves followed by one chroma-only CU over the node; every
    per-plane step gated by xevd_check_luma / xevd_check_chroma, maps written by the luma leaves only, HTDF on luma CUs only"""
    w, h, prm, cl, refs = dual_tree_inputs(variant, kw, bd, eipd, htdf, intra_frac)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    # the planes a dual-tree CU does not carry are untouched by it: dropping the chroma-only CUs changes chroma only
    keep = (cl.cus["flags"] & 3) != 2
    assert keep.sum() < len(cl.cus)


@pytest.mark.parametrize("kw,bd,intra_frac", [({}, 10, 1.0), (dict(log2_ctu=7), 8, 0.5), (dict(log2_ctu=5, suco=False), 10, 0.7)])
def test_recon_frame_dual_tree_ibc(oracle, reference, kw, bd, intra_frac):
    """luma-only leaves that are IBC CUs: xevdm_IBC_mc copies luma only under TREE_L (src_main/xevdm_mc.c:2059-2073), the chroma-only CU of
    the node then finds a non-intra SCU at the node's centre and falls back to IPD_DC as its DM mode (src_main/xevdm.c:1084-1091)"""
    w, h, prm, cl, refs = dual_tree_inputs("C", kw, bd, 1, 1, intra_frac, ibc=0.5)
    leaf_ibc = (cl.cus["mode"] == 4) & ((cl.cus["flags"] & 3) == 1)
    assert leaf_ibc.sum() >= 3, "test picture has no IBC leaves"
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("ibc", [0.0, 0.5])
@pytest.mark.parametrize("bd,addb", [(10, 0), (10, 1), (8, 0), (12, 1)])
def test_deblock_dual_tree(oracle, reference, bd, addb, ibc):
    """deblock_tree visits the luma leaves as TREE_L and then the node as TREE_C (src_main/xevdm.c:1991-1998): inner leaf edges are
    filtered in luma only, the node's outline in chroma once (xevdm_df.c:155-160,245-250,916-920,986-997)"""
    w, h, prm, cl, refs = dual_tree_inputs("C", {}, bd, 1, 0, 0.6, ibc=ibc)
    rng = np.random.default_rng(300 + bd + addb)
    prm.tool_addb = addb
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl in base.planes():
        pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
    if not ibc:
        synth.randomize_deblock_maps(base, cl, rng, intra_frac=0.15)       # else: the maps the reconstruction published (IBC / intra / cbf bits of the leaves)
    tbl, ids = synth.chroma_qp_table(True), ((0, 1), (1, 0))
    a = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    b = reference.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    changed = sum(int((x != y).sum()) for x, y in zip(a.planes(), base.planes()))
    assert changed > 300, "test picture does not exercise the filter"
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


def constrained_inputs(variant, kw, bd, eipd, htdf, intra_frac, dual):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=91, n_refs=2, coded_frac=0.7, **kw)
    prm.tool_eipd, prm.tool_htdf, prm.slice_qp, prm.constrained_intra_pred = eipd, htdf, 37, 1
    if dual:
        synth.split_local_dual_tree(cl, np.random.default_rng(5), 0.5)
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=bool(eipd), constrained=True)
    synth.derive_avail_cu(cl)
    cl.validate()
    return w, h, prm, cl, synth.make_refs(w, h, bd, 2, seed=9)


CONSTRAINED_CASES = [("C", {}, 10, 0, 0, 0.5, 0), ("C", {}, 10, 1, 1, 0.5, 0), ("C", dict(log2_ctu=5), 8, 1, 1, 0.3, 1), ("A", dict(log2_cu=3), 10, 1, 1, 0.6, 0),
                     ("C", dict(log2_ctu=7), 12, 1, 1, 0.5, 1), ("B", {}, 10, 0, 0, 0.4, 0)]


@pytest.mark.parametrize("variant,kw,bd,eipd,htdf,intra_frac,dual", CONSTRAINED_CASES)
def test_recon_frame_constrained_intra(oracle, reference, variant, kw, bd, eipd, htdf, intra_frac, dual):
    """pps.constrained_intra_pred_flag in pictures that mix intra and inter CUs: intra prediction takes neighbours from intra CUs only
    (xevd_get_nbr_b / xevdm_get_nbr with constrained_intra_flag, src_main/xevdm.c:609-652; the oracle gets that as the masks of
    synth.add_intra_cus) and so does the HTDF ring of an intra CU (xevdm_recon.c:317,338,359).  The reference applies its own tests
    on its own map_scu."""
    w, h, prm, cl, refs = constrained_inputs(variant, kw, bd, eipd, htdf, intra_frac, dual)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    prm.constrained_intra_pred = 0
    if htdf:
        c = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        assert (a.y != c.y).sum() > 50, "test picture does not exercise the constrained HTDF ring"


@pytest.mark.parametrize("seed,lg,addb", [(3, 6, 0), (4, 7, 0), (5, 5, 0), (3, 6, 1)])
def test_deblock_after_dmvr_and_affine(oracle, reference, seed, lg, addb):
    """which vectors the boundary strength compares once DMVR / affine have run: the ADDB walkers get mctx->map_unrefined_mv
    (src_main/xevdm.c:2009-2041), the Baseline-filter walkers read ctx->map_mv - refined sub-PU and affine sub-block vectors
    (src_main/xevdm_df.c:111-124,207-208,1143-1166).  Maps as the reconstruction published them."""
    w, h = 256, 136
    prm, cl, refs, alf, flags = synth.make_main_frame(w, h, bit_depth=10, seed=seed, log2_ctu=lg, iqt=bool(seed & 1))
    prm.tool_addb = addb
    pic = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    assert (pic.map_mv != pic.map_unrefined_mv).sum() > 50, "test picture has no refined / affine vectors"
    for pl in pic.planes():                      # small edge steps, so that the strength decides what the filter does
        pl[...] = (pl.astype(np.int32) // 8 + (1 << 9)).astype(np.int16)
    tbl, ids = synth.chroma_qp_table(True), ((0, 1), (1, 0))
    a = oracle.deblock_frame(prm, pic.copy(), cl, tbl, bool(addb), ids)
    b = reference.deblock_frame(prm, pic.copy(), cl, tbl, bool(addb), ids)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    if not addb and lg == 7:        # the choice is visible where an edge runs INSIDE a refined CU: the 64-sample transform edge of 128-sample CUs
        swapped = pic.copy()
        swapped.map_mv[...] = pic.map_unrefined_mv
        c = oracle.deblock_frame(prm, swapped, cl, tbl, False, ids)
        assert any((x != y).any() for x, y in zip(a.planes(), c.planes())), "the two vector maps give the same picture: the case pins nothing"


IBC_CASES = [("C", {}, 10, 0.3), ("C", dict(log2_ctu=7), 10, 0.3), ("C", dict(log2_ctu=5), 8, 0.5), ("B", {}, 10, 0.0), ("A", dict(log2_cu=3), 10, 0.2)]


def ibc_inputs(variant, kw, bd, intra_frac):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=51, n_refs=2, coded_frac=0.8, ats_inter_frac=0.3, **kw)
    prm.tool_eipd = 1
    prm.tool_ibc = 1
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True, ats_intra_frac=0.5, ibc_frac=0.6)
    cl.validate()
    assert (cl.cus["mode"] == 4).sum() >= 5
    return w, h, prm, cl, synth.make_refs(w, h, bd, 2, seed=9)


@pytest.mark.parametrize("variant,kw,bd,intra_frac", IBC_CASES)
def test_recon_frame_ibc(oracle, reference, variant, kw, bd, intra_frac):
    """intra block copy (xevdm_IBC_mc): whole-sample copies from the decoded part of the current picture, odd vectors included
    (chroma vector = luma >> 1), mixed with intra, inter and ats_inter CUs"""
    w, h, prm, cl, refs = ibc_inputs(variant, kw, bd, intra_frac)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


HTDF_CASES = [("C", {}, 10, 0.3, 32), ("C", dict(log2_ctu=7), 10, 0.3, 45), ("C", dict(log2_ctu=5), 8, 0.5, 22), ("B", {}, 10, 0.0, 37),
              ("A", dict(log2_cu=3), 10, 0.2, 27), ("B", {}, 10, 1.0, 17), ("B", {}, 12, 1.0, 51)]


def htdf_inputs(variant, kw, bd, intra_frac, qp):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=61, n_refs=2, coded_frac=0.7, ats_inter_frac=0.3, **kw)
    prm.tool_eipd = prm.tool_ibc = prm.tool_htdf = 1
    prm.slice_qp = qp
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True, ats_intra_frac=0.5, ibc_frac=0.3)
    synth.derive_avail_cu(cl)
    cl.validate()
    return w, h, prm, cl, synth.make_refs(w, h, bd, 2, seed=9)


@pytest.mark.parametrize("variant,kw,bd,intra_frac,qp", HTDF_CASES)
def test_recon_frame_htdf(oracle, reference, variant, kw, bd, intra_frac, qp):
    """Main tool_htdf: the in-loop Hadamard-domain luma filter after every intra CU and every CU with a luma residual, all five QP
    tables, the size / QP skip rules, ring samples from whichever neighbours xevd_get_avail_intra reports (right ones under SUCO)"""
    w, h, prm, cl, refs = htdf_inputs(variant, kw, bd, intra_frac, qp)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    if qp > 17:
        prm.tool_htdf = 0
        c = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        assert (a.y != c.y).sum() > 1000, "test picture does not exercise the filter"


DMVR_CASES = [("C", {}, 10, 3), ("C", dict(log2_ctu=7), 10, 0), ("C", dict(log2_ctu=5), 8, 2), ("B", {}, 10, 40), ("A", dict(log2_cu=3), 10, 1),
              ("B", dict(main_mv=True), 12, 6)]


@pytest.mark.parametrize("variant,kw,bd,noise", DMVR_CASES)
def test_recon_frame_dmvr(oracle, reference, variant, kw, bd, noise):
    """Main tool_dmvr (xevdm_mc / processDMVR): bilinear search planes, mirrored 5-point SAD search over two rounds, early exits,
    parabolic sub-sample step, final prediction from the replicated-border window, per 16x16 sub-PU; pictures AND the refined
    vectors published per SCU are compared"""
    w, h = 256, 136
    prm, cl, refs = synth.make_dmvr_case(w, h, bit_depth=bd, variant=variant, seed=71, noise=noise, coded_frac=0.5, **kw)
    cl.validate()
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    refined = ((a.map_scu >> 25) & 1).astype(bool)
    assert refined.sum() > 100 and (a.map_mv != a.map_unrefined_mv).any(axis=(1, 2)).sum() > 50, "test picture does not exercise DMVR"
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    inter = cl_inter_mask(cl, a.w_scu)
    assert np.array_equal(a.map_mv[inter], b.map_mv[inter])


def cl_inter_mask(cl, w_scu):
    """SCUs covered by inter CUs (the reference harness publishes vectors only for those)"""
    m = np.zeros(((cl.h + 3) >> 2, w_scu), bool)
    for cu in cl.cus:
        if int(cu["mode"]) == 1:
            m[int(cu["y"]) >> 2:(int(cu["y"]) >> 2) + (1 << (int(cu["log2h"]) - 2)), int(cu["x"]) >> 2:(int(cu["x"]) >> 2) + (1 << (int(cu["log2w"]) - 2))] = True
    return m.reshape(-1)


AFFINE_CASES = [("C", {}, 10), ("C", dict(log2_ctu=7), 10), ("C", dict(log2_ctu=5), 8), ("B", {}, 10), ("A", dict(log2_cu=3), 10), ("B", {}, 12),
                ("C", dict(mv_range_px=400), 10)]


def affine_inputs(variant, kw, bd):
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=81, n_refs=2, coded_frac=0.5, main_mv=True, ats_inter_frac=0.2, **kw)
    prm.tool_affine = 1
    synth.add_affine_cus(cl, np.random.default_rng(5), 0.7)
    cl.validate()
    return w, h, prm, cl, synth.make_refs(w, h, bd, 2, seed=9)


def cu_mask(cl, mode, w_scu):
    m = np.zeros(((cl.h + 3) >> 2, w_scu), bool)
    for cu in cl.cus[cl.cus["mode"] == mode]:
        m[int(cu["y"]) >> 2:(int(cu["y"]) >> 2) + (1 << (int(cu["log2h"]) - 2)), int(cu["x"]) >> 2:(int(cu["x"]) >> 2) + (1 << (int(cu["log2w"]) - 2))] = True
    return m.reshape(-1)


@pytest.mark.parametrize("variant,kw,bd", AFFINE_CASES)
def test_recon_frame_affine(oracle, reference, variant, kw, bd):
    """Main tool_affine (xevdm_affine_mc): 4- and 6-parameter models, the sub-block path (one vector per CU as the reference computes
    it), EIF with and without the clamped vector window, uni- and bi-prediction; pictures and the per-SCU vectors of
    xevdm_set_affine_mvf are compared"""
    w, h, prm, cl, refs = affine_inputs(variant, kw, bd)
    assert (cl.cus["mode"] == 5).sum() > 20
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))
    m = cu_mask(cl, 5, a.w_scu)
    assert np.array_equal(a.map_mv[m], b.map_mv[m])


def test_dra_apply(oracle, reference):
    """dynamic range adjustment on pull (xevd_apply_dra_chroma_plane / _luma_plane in the order xevd_apply_filter calls them)"""
    rng = np.random.default_rng(77)
    w, h = 192, 136
    dra = synth.make_dra_params(rng)
    pic = HostPicture.random(w, h, 10, rng)
    a = oracle.dra_apply(pic.copy(), dra)
    b = reference.dra_apply(pic.copy(), dra)
    assert (a.y != pic.y).sum() > 1000 and (a.u != pic.u).sum() > 1000
    for pa, pb, name in zip(a.planes(), b.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("tree,first", [(False, 0), (False, 2000), (True, 0), (True, 2000)])
def test_random_pipeline_vs_reference(reference, tree, first):
    """the randomised whole-pipeline cases the GPU is checked on (tests/test_gpu_fuzz.py: random size, bit depth, CTU size, partition,
    tool mix; with `tree` also local dual tree nodes and constrained intra prediction) - here the oracle against the reference itself:
    recon, deblocking, ALF, padding.  tools/ref_sweep.py runs the same over any seed range (2500 cases identical at the end of round 1)."""
    from tools.ref_sweep import sweep
    assert sweep(first, 40, tree) == []


def assert_maps_equal(a, b, what=""):
    """per-SCU maps of the oracle (a) against what the reference's own xevdm_set_dec_info published (b, oracle/ref_harness.c info_publish)"""
    for name in ("map_scu", "map_refi", "map_mv", "map_unrefined_mv"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x, y), f"{what}{name}: {int((x != y).sum())} entries differ, first SCU {int(np.argwhere((x != y).reshape(len(x), -1).any(axis=1))[0][0])}"


@pytest.mark.parametrize("case", ["inter_B", "inter_iqt_skipflags", "main_all_64", "main_all_128", "main_all_32", "dual_tree", "constrained_dual", "ibc", "ats"])
def test_set_dec_info_maps(oracle, reference, case):
    """xevd_set_dec_info / xevdm_set_dec_info (src_base/xevd_util.c:1574-1691, src_main/xevdm_util.c:4205-4389) pinned directly: the harness
    fills a zeroed XEVDM_CTX / XEVDM_CORE from every work item and calls the reference's function (also xevdm_set_affine_mvf and
    xevdm_set_cu_cbf_flags through it); map_scu (QP, intra / skip / cbf / DMVR / affine / IBC / COD bits), map_refi, map_mv (refined and
    affine sub-block vectors) and map_unrefined_mv must equal what the oracle - and therefore the GPU kernels - publish."""
    w, h = 256, 136
    if case == "inter_B":
        prm, cl = synth.make_inter_frame(w, h, bit_depth=10, variant="B", seed=41, n_refs=2, coded_frac=0.6)
        refs = synth.make_refs(w, h, 10, 2, seed=42)
    elif case == "inter_iqt_skipflags":
        prm, cl = synth.make_inter_frame(w, h, bit_depth=8, variant="C", seed=43, n_refs=2, coded_frac=0.5, iqt=True, main_mv=True)
        rng = np.random.default_rng(44)
        skip = (rng.random(cl.n_cu) < 0.3) & (cl.cus["cbf"] == 0)
        cl.cus["flags"] = np.where(skip, cl.cus["flags"] | 4, cl.cus["flags"])           # XB200_CUF_SKIP -> MCU_SET_SF
        cl.cus["qp_map"] = rng.integers(10, 50, cl.n_cu)
        refs = synth.make_refs(w, h, 8, 2, seed=45)
    elif case.startswith("main_all"):
        lg = {"64": 6, "128": 7, "32": 5}[case.split("_")[-1]]
        prm, cl, refs, _, _ = synth.make_main_frame(w, h, bit_depth=10, seed=46 + lg, log2_ctu=lg, iqt=lg != 7)
    elif case == "dual_tree":
        w, h, prm, cl, refs = dual_tree_inputs("C", dict(log2_ctu=6), 10, 1, 1, 0.7)
    elif case == "constrained_dual":
        w, h, prm, cl, refs = constrained_inputs("C", dict(log2_ctu=6), 10, 1, 1, 0.5, 1)
    elif case == "ibc":
        w, h, prm, cl, refs = ibc_inputs("C", dict(log2_ctu=6), 10, 0.3)
    else:
        prm, cl = synth.make_inter_frame(w, h, bit_depth=10, variant="C", seed=51, n_refs=2, coded_frac=0.8, iqt=True, ats_inter_frac=0.5)
        prm.tool_ats = 1
        refs = synth.make_refs(w, h, 10, 2, seed=52)
    a = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    b = reference.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pa, pb, n in zip(a.planes(), b.planes(), "YUV"):
        assert np.array_equal(pa, pb), n
    assert_maps_equal(a, b, case + ": ")
