"""GPU parity: picture-wide passes (deblocking, padding) through the C ABI against the CPU oracle."""
import numpy as np
import pytest

from xevd_b200 import abi, synth
from xevd_b200.frame import HostPicture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from xevd_b200.device import Context
    c = Context(0)
    yield c
    c.close()


def _smooth(pic, bd):
    for pl in pic.planes():
        pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)


@pytest.mark.parametrize("variant,log2_cu,bd,main_tbl", [("B", 4, 10, 0), ("A", 2, 10, 0), ("A", 3, 8, 1), ("B", 4, 8, 0), ("A", 6, 10, 1)])
def test_deblock_baseline(ctx, oracle, variant, log2_cu, bd, main_tbl):
    w, h = 192, 136
    rng = np.random.default_rng(40 + log2_cu + bd)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=31, n_refs=2, coded_frac=0.5, log2_cu=log2_cu,
                                     bi_frac=0.3, mv_range_px=3)
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    refs = synth.make_refs(w, h, bd, 2, seed=32)
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    _smooth(base, bd)
    synth.randomize_deblock_maps(base, cl, rng)
    tbl = synth.chroma_qp_table(bool(main_tbl))
    want = oracle.deblock_frame(prm, base.copy(), cl, tbl)

    d = ctx.pic_alloc(w, h).upload(base, padded=False).upload_maps(base, cl.edge_flags())
    ctx.set_chroma_qp_table(tbl)
    ctx.deblock(prm, d)
    got = d.download()
    d.free()
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


def test_recon_then_deblock_pipeline(ctx, oracle):
    """recon -> deblock -> pad on the device with the maps and edge flags the recon kernel itself published"""
    w, h, bd = 320, 200, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=51, n_refs=2, coded_frac=0.6, mv_range_px=24)
    cl.cus["qp_map"] = np.random.default_rng(3).integers(30, 46, cl.n_cu)
    refs = synth.make_refs(w, h, bd, 2, seed=52)
    for r in refs:                       # smooth references -> small edge steps -> the filter acts
        _smooth(r, bd)
        r.pad_borders()
    tbl = synth.chroma_qp_table(False)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    pre = want.copy()
    oracle.deblock_frame(prm, want, cl, tbl)
    oracle.pad(want)
    assert sum(int((x != y).sum()) for x, y in zip(want.planes(), pre.planes())) > 1000

    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    i = cur.info()
    ctx.deblock(prm, cur)
    ctx.pad(cur)
    got = cur.download_padded()
    for p in drefs + [cur]:
        p.free()
    assert np.array_equal(got.buf_y, want.buf_y) and np.array_equal(got.buf_u, want.buf_u) and np.array_equal(got.buf_v, want.buf_v)


@pytest.mark.parametrize("force", ["0", "1"])
def test_edge_map_published_by_recon(oracle, monkeypatch, force):
    """both recon kernels publish the CU/TU edge flags the deblocking pass consumes"""
    from xevd_b200.device import Context
    monkeypatch.setenv("XB200_FORCE_GENERIC", force)
    c = Context(0)
    w, h, bd = 256, 128, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=61, n_refs=1)
    refs = synth.make_refs(w, h, bd, 1, seed=62)
    drefs = [c.pic_alloc(w, h).upload(r) for r in refs]
    cur = c.pic_alloc(w, h)
    c.recon_frame(prm, cur, drefs, drefs, cl)
    edge = cur.download_edge_map()
    assert np.array_equal(edge, cl.edge_flags())
    c.close()


@pytest.mark.parametrize("variant,log2_cu,bd,aoff,boff", [("B", 4, 10, 0, 0), ("A", 2, 10, 2, -2), ("A", 3, 8, -3, 4), ("B", 4, 8, 6, 6), ("A", 6, 10, 0, 0), ("B", 4, 12, 0, 2)])
def test_deblock_addb(ctx, oracle, variant, log2_cu, bd, aoff, boff):
    """Main-profile deblocking (tool_addb) incl. reference-picture aliasing between indices and lists"""
    w, h = 192, 136
    rng = np.random.default_rng(140 + log2_cu + bd)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=33, n_refs=3, coded_frac=0.4, log2_cu=log2_cu,
                                     bi_frac=0.4, mv_range_px=2)
    prm.tool_addb = 1
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    prm.deblock_alpha_offset, prm.deblock_beta_offset = aoff, boff
    refs = synth.make_refs(w, h, bd, 3, seed=34)
    base = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    _smooth(base, bd)
    synth.randomize_deblock_maps(base, cl, rng, intra_frac=0.15)
    tbl = synth.chroma_qp_table(True)
    want = oracle.deblock_frame(prm, base.copy(), cl, tbl, True, ((0, 1, 0), (2, 1, 0)))

    pics = [ctx.pic_alloc(w, h) for _ in range(3)]          # three distinct reference pictures: ids 0, 1, 2
    d = ctx.pic_alloc(w, h).upload(base, padded=False).upload_maps(base, cl.edge_flags())
    ctx.set_chroma_qp_table(tbl)
    ctx.deblock(prm, d, [pics[0], pics[1], pics[0]], [pics[2], pics[1], pics[0]])
    got = d.download()
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for p in pics + [d]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


@pytest.mark.parametrize("w,h,bd,log2_ctu,enable", [(192, 136, 10, 6, (1, 1, 1)), (200, 72, 8, 6, (1, 0, 1)), (128, 128, 10, 5, (1, 1, 0)), (72, 200, 10, 6, (0, 1, 1)),
                                                    (256, 136, 10, 7, (1, 1, 1)), (1920, 1080, 10, 6, (1, 1, 1))])
def test_alf(ctx, oracle, w, h, bd, log2_ctu, enable):
    """adaptive loop filter (xevdm_alf.c) bit-exact: classes, transposes, mirrored picture borders, partial CTUs, CTU flags"""
    rng = np.random.default_rng(w * 3 + h + bd)
    p = HostPicture.random(w, h, bd, rng)
    yy, xx = np.mgrid[0:h, 0:w]
    p.y[...] = np.clip((p.y.astype(np.int32) >> 3) + ((xx * 3 + yy * 5) % 97) * (1 << (bd - 8)) + ((xx // 8 + yy // 8) % 2) * (40 << (bd - 8)), 0, (1 << bd) - 1).astype(np.int16)
    prm = abi.make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, enable)
    n_ctu = ((w + (1 << log2_ctu) - 1) >> log2_ctu) * ((h + (1 << log2_ctu) - 1) >> log2_ctu)
    flags = (rng.random(n_ctu) < 0.8).astype(np.uint8)
    want = oracle.alf_frame(prm, p.copy(), alf, flags)
    d = ctx.pic_alloc(w, h).upload(p, padded=False)
    ctx.alf(prm, d, alf, flags)
    got = d.download()
    for pa, pb, name in zip(got.planes(), want.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()), np.argwhere(pa != pb)[:5])
    # all CTUs on (NULL flags), twice in a row through the same context (scratch reuse)
    want2 = oracle.alf_frame(prm, want.copy(), alf, None)
    ctx.alf(prm, d, alf, None)
    got2 = d.download()
    for pa, pb, name in zip(got2.planes(), want2.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()))


@pytest.mark.parametrize("w,h,bd,log2_ctu,col_bd,row_bd", [(256, 128, 10, 6, [0, 2, 4], [0, 2]), (320, 192, 8, 5, [0, 3, 4, 10], [0, 2, 6]), (384, 256, 10, 7, [0, 1, 3], [0, 1, 2])])
def test_alf_tiles_without_filtering_across(ctx, oracle, w, h, bd, log2_ctu, col_bd, row_bd):
    """loop_filter_across_tiles_enabled_flag == 0: the ALF windows mirror at a tile border exactly as at the picture border
    (alf_process_tile, xevdm_alf.c:989-1046), so every tile must come out as if it were a picture of its own (the oracle on the cut-out)"""
    rng = np.random.default_rng(w + h + bd)
    p = HostPicture.random(w, h, bd, rng)
    prm = abi.make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, (1, 1, 1))
    ctu = 1 << log2_ctu
    w_ctu, h_ctu = (w + ctu - 1) >> log2_ctu, (h + ctu - 1) >> log2_ctu
    flags = (rng.random(w_ctu * h_ctu) < 0.8).astype(np.uint8)
    d = ctx.pic_alloc(w, h).upload(p, padded=False)
    ctx.set_tiles(col_bd, row_bd, across=False)
    try:
        ctx.alf(prm, d, alf, flags)
        got = d.download()
    finally:
        ctx.set_tiles()
        d.free()
    for tr in range(len(row_bd) - 1):
        for tc in range(len(col_bd) - 1):
            x0, x1, y0, y1 = col_bd[tc] * ctu, min(col_bd[tc + 1] * ctu, w), row_bd[tr] * ctu, min(row_bd[tr + 1] * ctu, h)
            sub = HostPicture(x1 - x0, y1 - y0, 0)
            sub.y[...] = p.y[y0:y1, x0:x1]; sub.u[...] = p.u[y0 // 2:y1 // 2, x0 // 2:x1 // 2]; sub.v[...] = p.v[y0 // 2:y1 // 2, x0 // 2:x1 // 2]
            sprm = abi.make_params(x1 - x0, y1 - y0, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
            sflags = np.ascontiguousarray(flags.reshape(h_ctu, w_ctu)[row_bd[tr]:row_bd[tr + 1], col_bd[tc]:col_bd[tc + 1]]).ravel()
            want = oracle.alf_frame(sprm, sub, alf, sflags)
            for a, b, sh, name in ((got.y, want.y, 0, "y"), (got.u, want.u, 1, "u"), (got.v, want.v, 1, "v")):
                cut = a[y0 >> sh:y1 >> sh, x0 >> sh:x1 >> sh]
                assert np.array_equal(cut, b), (tr, tc, name, int((cut != b).sum()), np.argwhere(cut != b)[:4].tolist())


def test_set_tiles_refuses_bad_grids(ctx):
    L, H = ctx.lib, ctx.handle
    ok = np.array([0, 2, 4], np.uint16)
    assert L.xb200_set_tiles(H, 2, ok.ctypes.data, 1, np.array([0, 9], np.uint16).ctypes.data, 0) == 0
    for cols in ([1, 2, 4], [0, 2, 2], [0, 3, 2]):
        bad = np.array(cols, np.uint16)
        assert L.xb200_set_tiles(H, 2, bad.ctypes.data, 1, np.array([0, 9], np.uint16).ctypes.data, 0) < 0
    assert L.xb200_set_tiles(H, 21, ok.ctypes.data, 1, ok.ctypes.data, 0) < 0
    assert L.xb200_set_tiles(H, 0, ok.ctypes.data, 1, ok.ctypes.data, 0) < 0
    ctx.set_tiles()


@pytest.mark.parametrize("kw,bd,addb", [({}, 10, 1), (dict(log2_ctu=7), 10, 1), (dict(log2_ctu=5), 8, 1), ({}, 10, 0), (dict(log2_ctu=7), 8, 0)])
def test_deblock_main_partitions(ctx, oracle, kw, bd, addb):
    """both deblocking filters on Main-profile partitions (non-square CUs, ternary-split edges, 128-sample CUs, ats_inter CUs)"""
    from tests.test_oracle_vs_ref import deblock_main_inputs
    w, h, prm, cl, base, tbl, ids = deblock_main_inputs(oracle, kw, bd, addb)
    want = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    pics = [ctx.pic_alloc(w, h) for _ in range(3)]
    d = ctx.pic_alloc(w, h).upload(base, padded=False).upload_maps(base, cl.edge_flags())
    ctx.set_chroma_qp_table(tbl)
    ctx.deblock(prm, d, [pics[0], pics[1], pics[0]], [pics[2], pics[1], pics[0]])
    got = d.download()
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for p in pics + [d]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


@pytest.mark.parametrize("across", [0, 1])
@pytest.mark.parametrize("kw,bd,addb,grid", [({}, 10, 1, (2, 2)), (dict(log2_ctu=5), 8, 1, (3, 2)), ({}, 10, 0, (2, 2)), (dict(log2_ctu=5), 10, 0, (4, 3)),
                                             (dict(log2_ctu=7), 8, 0, (2, 1))])
def test_deblock_tiles(ctx, oracle, kw, bd, addb, grid, across):
    """pictures of several tiles (xb200_set_tiles): tile-boundary edges are filtered only with loop_filter_across_tiles_enabled_flag;
    the oracle's rule is pinned to the reference's deblock functions in tests/test_oracle_vs_ref.py::test_deblock_tiles"""
    from tests.test_oracle_vs_ref import deblock_main_inputs, tile_grid
    w, h, prm, cl, base, tbl, ids = deblock_main_inputs(oracle, kw, bd, addb)
    cb, rb = tile_grid(w, h, prm.log2_ctu, *grid)
    oracle.set_tiles(cb, rb, bool(across))
    try:
        want = oracle.deblock_frame(prm, base.copy(), cl, tbl, bool(addb), ids)
    finally:
        oracle.set_tiles()
    pics = [ctx.pic_alloc(w, h) for _ in range(3)]
    d = ctx.pic_alloc(w, h).upload(base, padded=False).upload_maps(base, cl.edge_flags())
    ctx.set_chroma_qp_table(tbl)
    ctx.set_tiles(cb, rb, across=bool(across))
    try:
        ctx.deblock(prm, d, [pics[0], pics[1], pics[0]], [pics[2], pics[1], pics[0]])
        got = d.download()
    finally:
        ctx.set_tiles()
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in pics + [d]:
            p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


@pytest.mark.parametrize("across", [0, 1])
@pytest.mark.parametrize("w,h,bd,log2_ctu,grid", [(256, 136, 10, 6, (2, 2)), (200, 120, 8, 5, (3, 2)), (384, 256, 10, 7, (3, 1)), (320, 192, 10, 5, (4, 4)),
                                                  (1920, 1080, 10, 6, (4, 3))])
def test_alf_tiles(ctx, oracle, w, h, bd, log2_ctu, grid, across):
    """ALF on pictures of several tiles, both values of the flag, against the oracle (pinned to alf_process_tile in
    tests/test_oracle_vs_ref.py::test_alf_tiles)"""
    from tests.test_oracle_vs_ref import tile_grid
    rng = np.random.default_rng(w + h + bd + across)
    p = HostPicture.random(w, h, bd, rng)
    prm = abi.make_params(w, h, bit_depth=bd, log2_ctu=log2_ctu, tool_alf=1)
    alf = synth.make_alf_params(rng, (1, 1, 1))
    n_ctu = ((w + (1 << log2_ctu) - 1) >> log2_ctu) * ((h + (1 << log2_ctu) - 1) >> log2_ctu)
    flags = (rng.random(n_ctu) < 0.8).astype(np.uint8)
    cb, rb = tile_grid(w, h, log2_ctu, *grid)
    oracle.set_tiles(cb, rb, bool(across))
    try:
        want = oracle.alf_frame(prm, p.copy(), alf, flags)
    finally:
        oracle.set_tiles()
    d = ctx.pic_alloc(w, h).upload(p, padded=False)
    ctx.set_tiles(cb, rb, across=bool(across))
    try:
        ctx.alf(prm, d, alf, flags)
        got = d.download()
    finally:
        ctx.set_tiles()
        d.free()
    for pa, pb, name in zip(got.planes(), want.planes(), "yuv"):
        assert np.array_equal(pa, pb), (name, int((pa != pb).sum()), np.argwhere(pa != pb)[:5].tolist())


@pytest.mark.parametrize("bd,addb,kw", [(10, 0, {}), (8, 0, dict(log2_ctu=5)), (10, 0, dict(log2_ctu=7)), (10, 1, {}), (10, 0, dict(suco=False))])
def test_deblock_suco_order_4wide(ctx, oracle, bd, addb, kw):
    """recon -> deblock on binary/ternary partitions down to 4-wide CUs in SUCO order: neighbouring chroma edges are 2 samples apart, each
    reads a sample the other writes, and the reference filters an edge when the LATER of its two CUs is visited (xevdm_df.c:272-300), so the
    run is not walked left to right (found on generated Main streams).  The recon kernels publish the order (tool_suco)."""
    w, h = 320, 192
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=77, n_refs=2, coded_frac=0.5, bi_frac=0.3, min_log2=2, **kw)
    prm.tool_addb = addb
    prm.qp_u_offset, prm.qp_v_offset = 3, -2
    cl.cus["qp_map"] = 44
    tbl, ids = synth.chroma_qp_table(True), ((0, 1), (1, 0))
    refs = synth.make_refs(w, h, bd, 2, seed=78)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    rec = want.copy()
    oracle.deblock_frame(prm, want, cl, tbl, bool(addb), ids)
    assert sum(int((a != b).sum()) for a, b in zip(want.planes()[1:], rec.planes()[1:])) > (20 if addb else 500), "chroma deblocking does nothing here"
    ctx.set_chroma_qp_table(tbl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    ctx.deblock(prm, cur, drefs, drefs[::-1])
    got = cur.download()
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


@pytest.mark.parametrize("bd,addb,kw", [(10, 0, {}), (10, 1, {}), (8, 0, dict(log2_ctu=5)), (12, 1, dict(log2_ctu=7))])
def test_dual_tree_pipeline(ctx, oracle, bd, addb, kw):
    """recon -> deblock -> pad of a picture with local dual tree nodes: the deblocking pass runs on the maps and the edge map the recon
    kernels left on the device (inner leaf edges filtered in luma only, the node's outline in chroma; xevdm_df.c:155-160,245-250)"""
    from tests.test_oracle_vs_ref import dual_tree_inputs
    w, h, prm, cl, refs = dual_tree_inputs("C", kw, bd, 1, 0, 0.6)
    prm.tool_addb = addb
    prm.qp_u_offset, prm.qp_v_offset = 2, -3
    tbl, ids = synth.chroma_qp_table(True), ((0, 1), (1, 0))
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    pre = want.copy()
    oracle.deblock_frame(prm, want, cl, tbl, bool(addb), ids)
    oracle.pad(want)
    assert sum(int((x != y).sum()) for x, y in zip(want.planes(), pre.planes())) > 300
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.set_chroma_qp_table(tbl)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    ctx.deblock(prm, cur, drefs, drefs[::-1])
    ctx.pad(cur)
    got = cur.download_padded()
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for p in drefs + [cur]:
        p.free()
    assert np.array_equal(got.buf_y, want.buf_y) and np.array_equal(got.buf_u, want.buf_u) and np.array_equal(got.buf_v, want.buf_v)


@pytest.mark.parametrize("use_dra,out_bits,crop", [(0, 16, (0, 0, 0, 0)), (1, 16, (0, 0, 0, 0)), (0, 8, (8, 16, 2, 6)), (1, 8, (4, 0, 0, 10)), (1, 16, (2, 2, 2, 2))])
def test_output_path(ctx, oracle, use_dra, out_bits, crop):
    """SURVEY 8f N2 + N4: what xevd_pull hands out - DRA on a copy, SPS cropping window, optional 16 -> 8-bit conversion - in one kernel"""
    from oracle.pyoracle import oracle_output
    rng = np.random.default_rng(5 + use_dra + out_bits)
    w, h = 200, 136
    pic = HostPicture.random(w, h, 10, rng)
    dra = synth.make_dra_params(rng) if use_dra else None
    d = ctx.pic_alloc(w, h).upload(pic, padded=False)
    got = d.pull(dra, out_bits, crop)
    still = d.download()
    d.free()
    ref = pic.copy()
    if use_dra:
        oracle.dra_apply(ref, dra)
    want = oracle_output(oracle, ref, out_bits, crop)
    for a, b, n in zip(got, want, "YUV"):
        assert a.shape == b.shape and np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    for a, b in zip(still.planes(), pic.planes()):
        assert np.array_equal(a, b), "the device picture must stay unfiltered (it is a reference picture)"
