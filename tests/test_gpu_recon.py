"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs."""
import numpy as np
import pytest

from xevd_b200 import synth
from xevd_b200.frame import HostPicture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from xevd_b200.device import Context
    c = Context(0)
    yield c
    c.close()


def _run(ctx, oracle, w, h, bd, variant, seed, n_refs=2, **kw):
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=seed, n_refs=n_refs, **kw)
    cl.validate()
    refs = synth.make_refs(w, h, bd, n_refs, seed=seed + 100)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_mv, want.map_mv)
    assert np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(got.map_scu, want.map_scu)


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("bd", [8, 10])
def test_inter_small(ctx, oracle, variant, bd):
    _run(ctx, oracle, 256, 136, bd, variant, seed=3, coded_frac=0.8)


def test_inter_large_mv_clipping(ctx, oracle):
    # vectors far outside the picture: exercises xevd_mv_clip and the T3 variant/phase split
    _run(ctx, oracle, 192, 128, 10, "B", seed=4, mv_range_px=400)


def test_inter_iqt(ctx, oracle):
    _run(ctx, oracle, 256, 128, 10, "B", seed=5, iqt=True)


def test_inter_main_tables(ctx, oracle):
    _run(ctx, oracle, 256, 128, 10, "B", seed=6, main_mv=True)


def test_inter_8x8_cus(ctx, oracle):
    _run(ctx, oracle, 128, 64, 10, "A", seed=7, log2_cu=3)


def test_inter_1080p(ctx, oracle):
    _run(ctx, oracle, 1920, 1080, 10, "A", seed=8, n_refs=1)


def test_pad(ctx, oracle):
    rng = np.random.default_rng(1)
    p = HostPicture.random(200, 104, 10, rng)
    d = ctx.pic_alloc(200, 104).upload(p, padded=True)
    got = d.download_padded()
    want = oracle.pad(p.copy())
    assert np.array_equal(got.buf_y, want.buf_y) and np.array_equal(got.buf_u, want.buf_u) and np.array_equal(got.buf_v, want.buf_v)
    d.free()


def test_fast_and_generic_kernels_agree(oracle, monkeypatch):
    """the throughput kernel (xb_recon2.cuh) and the generic kernel (xb_recon.cuh) are two implementations of the
    same contract: run both on one picture (XB200_FORCE_GENERIC selects the generic one at context creation)"""
    from xevd_b200.device import Context
    w, h, bd = 320, 192, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=21, n_refs=2, coded_frac=0.7)
    refs = synth.make_refs(w, h, bd, 2, seed=22)
    outs = []
    for force in ("0", "1"):
        monkeypatch.setenv("XB200_FORCE_GENERIC", force)
        c = Context(0)
        drefs = [c.pic_alloc(w, h).upload(r) for r in refs]
        cur = c.pic_alloc(w, h)
        c.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        outs.append(cur.download(maps=True))
        c.close()
    for a, b in zip(outs[0].planes(), outs[1].planes()):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0].map_mv, outs[1].map_mv) and np.array_equal(outs[0].map_scu, outs[1].map_scu)


def test_inter_4x4_cus(ctx, oracle):
    # smallest Baseline CUs: 4x4 luma with 2x2 chroma transform blocks, multiple MC rounds per CTU
    _run(ctx, oracle, 64, 64, 10, "A", seed=9, log2_cu=2)


def test_inter_64x64_cus(ctx, oracle):
    _run(ctx, oracle, 256, 128, 10, "A", seed=10, log2_cu=6, n_refs=2, bi_frac=0.5)


def test_recon_saturating_residual(ctx, oracle):
    # huge coefficients: dequant clips to s16, the residual add wraps to 16 bits before the pixel clip (xevd_recon.c:60)
    w, h, bd = 128, 64, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="A", seed=12, n_refs=1)
    rng = np.random.default_rng(5)
    idx = rng.integers(0, cl.coef.size, 200)
    cl.coef[idx] = rng.choice(np.array([-32768, 32767, -20000, 20000], np.int16), 200)
    cl.coef[cl.cus["coef_off"][::7]] = 3000      # large DC terms: residuals beyond +-2^15 before the clip
    refs = synth.make_refs(w, h, bd, 1, seed=13)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs, cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs, cl)
    got = cur.download()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


def test_reference_index_outside_lists_is_rejected(ctx):
    """bi-predicted CUs but an empty list 1: the host entry point refuses instead of launching"""
    from xevd_b200.device import XevdB200Error
    from xevd_b200 import abi
    w, h, bd = 128, 64, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=71, n_refs=1)
    assert (cl.cus["refi"][:, 1] >= 0).any()
    refs = synth.make_refs(w, h, bd, 1, seed=72)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    with pytest.raises(XevdB200Error) as e:
        ctx.recon_frame(prm, cur, drefs, [], cl)
    assert e.value.code == abi.XB200_ERR_INVALID_ARGUMENT
    for p in drefs + [cur]:
        p.free()


@pytest.mark.parametrize("variant,kw,bd,eipd,htdf,intra_frac", [("C", {}, 10, 0, 0, 1.0), ("C", {}, 10, 1, 0, 1.0), ("C", dict(log2_ctu=5), 8, 1, 1, 1.0),
                                                                ("A", dict(log2_cu=3), 10, 1, 1, 1.0), ("C", dict(suco=False), 12, 0, 0, 0.5),
                                                                ("C", dict(log2_ctu=7), 10, 1, 1, 0.4), ("C", dict(iqt=True), 10, 1, 0, 0.3)])
@pytest.mark.parametrize("force", ["0", "1"])
def test_dual_tree(oracle, monkeypatch, variant, kw, bd, eipd, htdf, intra_frac, force):
    """local dual tree (src_main/xevdm.c:1828-1846): luma-only leaves + one chroma-only CU per node.  Planes, per-SCU maps (published by
    the luma leaves only) and the edge map (inner leaf edges luma-only) against the oracle, through the per-CU dispatch (64x64 CTUs:
    dual-tree CUs in the generic kernel, the rest in the throughput kernel) and through the generic kernel alone"""
    from xevd_b200.device import Context
    from tests.test_oracle_vs_ref import dual_tree_inputs
    monkeypatch.setenv("XB200_FORCE_GENERIC", force)
    c = Context(0)
    w, h, prm, cl, refs = dual_tree_inputs(variant, kw, bd, eipd, htdf, intra_frac)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [c.pic_alloc(w, h).upload(r) for r in refs]
    cur = c.pic_alloc(w, h)
    c.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    edge = cur.download_edge_map()
    c.close()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(edge, cl.edge_flags())


@pytest.mark.parametrize("variant,kw,bd,eipd,htdf,intra_frac,dual", [("C", {}, 10, 0, 0, 0.5, 0), ("C", {}, 10, 1, 1, 0.5, 0), ("C", dict(log2_ctu=5), 8, 1, 1, 0.3, 1),
                                                                     ("A", dict(log2_cu=3), 10, 1, 1, 0.6, 0), ("C", dict(log2_ctu=7), 12, 1, 1, 0.5, 1),
                                                                     ("B", {}, 10, 0, 0, 0.4, 0)])
def test_constrained_intra(ctx, oracle, variant, kw, bd, eipd, htdf, intra_frac, dual):
    """pps.constrained_intra_pred_flag in mixed pictures: neighbour masks with the intra test and the HTDF ring of intra CUs taking left /
    right / upper samples from intra neighbours only (xevdm_recon.c:317,338,359; read from the published map_scu on the device)"""
    from tests.test_oracle_vs_ref import constrained_inputs
    w, h, prm, cl, refs = constrained_inputs(variant, kw, bd, eipd, htdf, intra_frac, dual)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("kw,bd,intra_frac", [({}, 10, 1.0), (dict(log2_ctu=7), 8, 0.5), (dict(log2_ctu=5, suco=False), 10, 0.7)])
def test_dual_tree_ibc(ctx, oracle, kw, bd, intra_frac):
    """luma-only leaves that are IBC CUs (luma copy only), next to intra leaves and the chroma-only CU of the node"""
    from tests.test_oracle_vs_ref import dual_tree_inputs
    w, h, prm, cl, refs = dual_tree_inputs("C", kw, bd, 1, 1, intra_frac, ibc=0.5)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    edge = cur.download_edge_map()
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(edge, cl.edge_flags())


@pytest.mark.parametrize("intra_frac,lg", [(0.08, 6), (1.0, 6), (0.3, 7)])
def test_dual_tree_1080p_pipeline(ctx, oracle, intra_frac, lg):
    """a full-size Main picture (IQT, 1/16-pel motion, EIPD, HTDF, constrained intra, ADDB) with local dual tree nodes scattered over
    500+ CTUs: sparse wavefront dependencies, per-CU dispatch between the kernels, recon -> deblock -> pad.  Decoded three times into
    the same picture: a missed dependency is a race."""
    w, h, bd = 1920, 1080, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=101, n_refs=2, coded_frac=0.4, main_mv=True, iqt=True, log2_ctu=lg)
    prm.tool_eipd = prm.tool_htdf = prm.tool_addb = prm.constrained_intra_pred = 1
    prm.slice_qp = 35
    synth.split_local_dual_tree(cl, np.random.default_rng(5), 0.7)
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True, constrained=True)
    synth.derive_avail_cu(cl)
    cl.validate()
    assert ((cl.cus["flags"] & 3) != 3).sum() > 200
    refs = synth.make_refs(w, h, bd, 2, seed=9)
    tbl = synth.chroma_qp_table(True)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    rec = want.copy()
    oracle.deblock_frame(prm, want, cl, tbl, True, ((0, 1), (1, 0)))
    oracle.pad(want)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.set_chroma_qp_table(tbl)
    try:
        for it in range(3):
            ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
            got = cur.download(maps=True)
            for a, b, n in zip(got.planes(), rec.planes(), "YUV"):
                assert np.array_equal(a, b), f"pass {it} recon plane {n}: {int((a != b).sum())} samples differ"
            assert np.array_equal(got.map_scu, rec.map_scu)
            ctx.deblock(prm, cur, drefs, drefs[::-1])
            ctx.pad(cur)
            out = cur.download_padded()
            assert np.array_equal(out.buf_y, want.buf_y) and np.array_equal(out.buf_u, want.buf_u) and np.array_equal(out.buf_v, want.buf_v), f"pass {it}"
    finally:
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in drefs + [cur]:
            p.free()


def test_dual_tree_inter_cu_is_refused(ctx):
    """an inter CU is always TREE_LC (xevdm.c:1122): one flagged luma-only is a caller error, not something to reconstruct"""
    from xevd_b200.device import XevdB200Error
    from xevd_b200 import abi
    w, h, bd = 128, 64, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="A", seed=73, n_refs=1)
    cl.cus["flags"][3] = (int(cl.cus["flags"][3]) & ~3) | 1
    refs = synth.make_refs(w, h, bd, 1, seed=74)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    with pytest.raises(XevdB200Error) as e:
        ctx.recon_frame(prm, cur, drefs, drefs, cl)
    assert e.value.code == abi.XB200_ERR_INVALID_ARGUMENT
    for p in drefs + [cur]:
        p.free()


@pytest.mark.parametrize("variant,log2_cu,bd,intra_frac", [("B", 4, 10, 1.0), ("A", 2, 10, 1.0), ("B", 4, 8, 0.4), ("A", 6, 10, 1.0), ("A", 3, 10, 0.5)])
@pytest.mark.parametrize("force", ["0", "1"])
def test_intra_baseline(oracle, monkeypatch, variant, log2_cu, bd, intra_frac, force):
    """I pictures and mixed pictures: inter CUs by the parallel kernel, intra CUs by the CTU wavefront kernel"""
    from xevd_b200.device import Context
    monkeypatch.setenv("XB200_FORCE_GENERIC", force)
    c = Context(0)
    w, h = 200, 136
    rng = np.random.default_rng(7)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=81, n_refs=2, coded_frac=0.7, log2_cu=log2_cu)
    synth.add_intra_cus(cl, rng, intra_frac)
    refs = synth.make_refs(w, h, bd, 2, seed=82)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [c.pic_alloc(w, h).upload(r) for r in refs]
    cur = c.pic_alloc(w, h)
    c.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    c.close()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("kw,bd", [({}, 10), (dict(log2_ctu=7), 10), (dict(log2_ctu=5), 8), (dict(suco=False), 10), (dict(iqt=True), 10)])
def test_inter_btt(ctx, oracle, kw, bd):
    """Main-profile partitions: non-square CUs, 32/64/128 CTUs, 128-sample CUs with gated 64x64 transform sub-blocks"""
    _run(ctx, oracle, 256, 136, bd, "C", seed=31, coded_frac=0.8, **kw)


@pytest.mark.parametrize("variant,kw,bd,intra_frac", [("C", {}, 10, 1.0), ("C", dict(log2_ctu=7), 10, 1.0), ("C", dict(log2_ctu=5), 8, 0.5),
                                                      ("B", {}, 10, 1.0), ("A", dict(log2_cu=2), 10, 1.0), ("C", dict(suco=False), 12, 0.7),
                                                      ("C", dict(iqt=True), 10, 0.6)])
def test_intra_eipd(ctx, oracle, variant, kw, bd, intra_frac):
    """Main-profile intra (tool_eipd): 33 luma / 5 chroma modes, left / up / right reference arrays, every avail_lr case"""
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=21, n_refs=2, coded_frac=0.7, **kw)
    prm.tool_eipd = 1
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True)
    refs = synth.make_refs(w, h, bd, 2, seed=9)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("variant,kw,bd,intra_frac,iqt", [("C", {}, 10, 0.5, 0), ("C", dict(log2_ctu=7), 10, 0.3, 1), ("C", dict(log2_ctu=5), 8, 0.5, 0),
                                                          ("B", {}, 10, 0.0, 1), ("A", dict(log2_cu=3), 10, 1.0, 0)])
def test_ats(ctx, oracle, variant, kw, bd, intra_frac, iqt):
    """Main tool_ats: DST-7 / DCT-8 luma transforms of intra CUs and sub-block transforms (half / quarter TU) of inter CUs"""
    w, h = 256, 136
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=41, n_refs=2, coded_frac=0.8, ats_inter_frac=0.6, iqt=bool(iqt), **kw)
    prm.tool_eipd = 1
    synth.add_intra_cus(cl, np.random.default_rng(4), intra_frac, eipd=True, ats_intra_frac=0.7)
    refs = synth.make_refs(w, h, bd, 2, seed=9)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("variant,kw,bd,intra_frac", [("C", {}, 10, 0.3), ("C", dict(log2_ctu=7), 10, 0.3), ("C", dict(log2_ctu=5), 8, 0.5), ("B", {}, 10, 0.0),
                                                      ("A", dict(log2_cu=3), 10, 0.2)])
def test_ibc(ctx, oracle, variant, kw, bd, intra_frac):
    """intra block copy CUs in the wavefront kernel, mixed with intra, inter and ats_inter CUs"""
    from tests.test_oracle_vs_ref import ibc_inputs
    w, h, prm, cl, refs = ibc_inputs(variant, kw, bd, intra_frac)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("variant,kw,bd,intra_frac,qp", [("C", {}, 10, 0.3, 32), ("C", dict(log2_ctu=7), 10, 0.3, 45), ("C", dict(log2_ctu=5), 8, 0.5, 22),
                                                         ("B", {}, 10, 0.0, 37), ("A", dict(log2_cu=3), 10, 0.2, 27), ("B", {}, 10, 1.0, 17), ("B", {}, 12, 1.0, 51)])
def test_htdf(ctx, oracle, variant, kw, bd, intra_frac, qp):
    """Main tool_htdf: in-order luma post-filter in the wavefront kernel (intra CUs and inter CUs with a luma residual)"""
    from tests.test_oracle_vs_ref import htdf_inputs
    w, h, prm, cl, refs = htdf_inputs(variant, kw, bd, intra_frac, qp)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)


@pytest.mark.parametrize("variant,kw,bd,noise", [("C", {}, 10, 3), ("C", dict(log2_ctu=7), 10, 0), ("C", dict(log2_ctu=5), 8, 2), ("B", {}, 10, 40),
                                                 ("A", dict(log2_cu=3), 10, 1), ("B", dict(main_mv=True), 12, 6), ("B", dict(iqt=True), 10, 2)])
def test_dmvr(ctx, oracle, variant, kw, bd, noise):
    """Main tool_dmvr: per-sub-PU refinement in the inter kernel; planes, refined map_mv, map_unrefined_mv and the DMVR flag bit"""
    w, h = 256, 136
    prm, cl, refs = synth.make_dmvr_case(w, h, bit_depth=bd, variant=variant, seed=71, noise=noise, coded_frac=0.5, **kw)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(got.map_unrefined_mv, want.map_unrefined_mv)


@pytest.mark.parametrize("variant,kw,bd", [("C", {}, 10), ("C", dict(log2_ctu=7), 10), ("C", dict(log2_ctu=5), 8), ("B", {}, 10), ("A", dict(log2_cu=3), 10),
                                           ("B", {}, 12), ("C", dict(mv_range_px=400), 10), ("B", dict(iqt=True), 10)])
def test_affine(ctx, oracle, variant, kw, bd):
    """Main tool_affine: sub-block and EIF prediction in the inter kernel, per-SCU model vectors in map_mv"""
    from tests.test_oracle_vs_ref import affine_inputs
    w, h, prm, cl, refs = affine_inputs(variant, kw, bd)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(got.map_unrefined_mv, want.map_unrefined_mv)


def test_intra_1080p_wavefront(ctx, oracle):
    """a full-size I picture: 510 CTUs through the wavefront (ticket + done flags)"""
    w, h, bd = 1920, 1080, 10
    rng = np.random.default_rng(9)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="A", seed=91, n_refs=1, coded_frac=0.5)
    synth.add_intra_cus(cl, rng, 1.0)
    refs = synth.make_refs(w, h, bd, 1, seed=92)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs, cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs, cl)
    got = cur.download()
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


@pytest.mark.gpu
@pytest.mark.parametrize("coded,intra,kw,qp", [(0.1, 0.02, {}, 32), (0.3, 0.05, {}, 40), (0.6, 0.1, {}, 27), (0.15, 0.03, dict(log2_ctu=7), 35), (0.2, 0.0, dict(log2_ctu=5), 30)])
def test_htdf_sparse_wavefront(ctx, oracle, coded, intra, kw, qp):
    """HTDF without IBC: CTUs wait only for neighbours whose filtered / intra CUs lie under the samples their own border CUs read; uncoded
    CUs break the chain.  Decoded several times: a missed dependency is a race."""
    w, h, bd = 1280, 712, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=71, n_refs=2, coded_frac=coded, **kw)
    prm.tool_eipd = prm.tool_htdf = 1
    prm.slice_qp = qp
    if intra:
        synth.add_intra_cus(cl, np.random.default_rng(4), intra, eipd=True)
    synth.derive_avail_cu(cl)
    cl.validate()
    refs = synth.make_refs(w, h, bd, 2, seed=72)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    try:
        for rep in range(4):
            ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
            got = cur.download()
            for a, b, n in zip(got.planes(), want.planes(), "YUV"):
                assert np.array_equal(a, b), f"repetition {rep}, plane {n}: {int((a != b).sum())} samples differ"
    finally:
        for p in drefs + [cur]:
            p.free()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,iqt,aff,ats,dmvr", [("C", 1, 0.03, 0.02, 0), ("B", 1, 0.05, 0.0, 0), ("C", 0, 0.0, 0.03, 0), ("C", 1, 0.02, 0.02, 1)])
def test_main_picture_per_ctu_dispatch(oracle, monkeypatch, variant, iqt, aff, ats, dmvr):
    """Main pictures with only a few ATS / affine / DMVR CUs: the throughput kernel reconstructs the CTUs without them, the generic kernel
    the others.  Planes and every per-SCU map must equal the oracle's, and the all-generic route must agree too."""
    from xevd_b200.device import Context
    w, h, bd = 960, 520, 10
    if dmvr:
        # DMVR on a minority of the CUs so that some CTUs stay with the throughput kernel
        prm, cl, refs = synth.make_dmvr_case(w, h, bit_depth=bd, variant=variant, seed=33, flag_frac=0.03, coded_frac=0.6, main_mv=True, ats_inter_frac=ats,
                                             iqt=bool(iqt))
    else:
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=33, n_refs=2, coded_frac=0.6, main_mv=True, ats_inter_frac=ats, iqt=bool(iqt))
        refs = synth.make_refs(w, h, bd, 2, seed=34)
    if aff:
        prm.tool_affine = 1
        synth.add_affine_cus(cl, np.random.default_rng(5), aff)
    cl.validate()
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    got = {}
    for force in ("0", "1"):
        monkeypatch.setenv("XB200_FORCE_GENERIC", force)
        c = Context(0)
        drefs = [c.pic_alloc(w, h).upload(r) for r in refs]
        cur = c.pic_alloc(w, h)
        c.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got[force] = (cur.download(maps=True), c.launches)
        c.close()
    assert got["0"][1] == got["1"][1] + 1, "per-CTU dispatch is two inter launches, the forced generic route one"
    for force in ("0", "1"):
        g = got[force][0]
        for a, b, n in zip(g.planes(), want.planes(), "YUV"):
            assert np.array_equal(a, b), f"force={force} plane {n}: {int((a != b).sum())} samples differ"
        assert np.array_equal(g.map_scu, want.map_scu) and np.array_equal(g.map_mv, want.map_mv) and np.array_equal(g.map_refi, want.map_refi)
        assert np.array_equal(g.map_unrefined_mv, want.map_unrefined_mv)


@pytest.mark.gpu
@pytest.mark.parametrize("frac,eipd,kw", [(0.05, 0, {}), (0.2, 0, {}), (0.5, 0, {}), (0.2, 1, {}), (0.2, 1, dict(log2_ctu=7)), (0.3, 0, dict(log2_ctu=5))])
def test_mixed_picture_sparse_wavefront(ctx, oracle, frac, eipd, kw):
    """P/B pictures with scattered intra CUs: without HTDF and IBC a CTU waits only for the neighbour CTUs whose intra CUs lie under the
    reference samples of its own border CUs.  A missed dependency is a race, so the picture is decoded several times."""
    w, h, bd = 1280, 712, 10
    rng = np.random.default_rng(17)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=61, n_refs=2, coded_frac=0.6, **kw)
    prm.tool_eipd = eipd
    synth.add_intra_cus(cl, rng, frac, eipd=bool(eipd))
    synth.derive_avail_cu(cl)
    refs = synth.make_refs(w, h, bd, 2, seed=62)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    try:
        for rep in range(4):
            ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
            got = cur.download()
            for a, b, n in zip(got.planes(), want.planes(), "YUV"):
                assert np.array_equal(a, b), f"repetition {rep}, plane {n}: {int((a != b).sum())} samples differ"
    finally:
        for p in drefs + [cur]:
            p.free()


def test_golden_frames_gpu(ctx):
    """the CUDA path against the committed golden vectors of the unmodified reference (tests/golden/frames.npz):
    recon, then deblocking + padding of the same picture"""
    from pathlib import Path
    from tests.test_golden import FRAME_CFGS, golden_frame_inputs
    z = np.load(Path(__file__).resolve().parent / "golden" / "frames.npz")
    for name, kw, intra in FRAME_CFGS:
        w, h, prm, cl, refs = golden_frame_inputs(name, kw, intra)
        bd = kw["bit_depth"]
        drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
        cur = ctx.pic_alloc(w, h)
        ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got = cur.download(maps=True)
        for pl, k in zip(got.planes(), "yuv"):
            assert np.array_equal(pl, z[f"{name}_{k}"]), (name, k)
        if not kw.get("iqt"):
            for pl in got.planes():
                pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
            synth.randomize_deblock_maps(got, cl, np.random.default_rng(8))
            cur.upload(got, padded=False).upload_maps(got)
            ctx.deblock(prm, cur)
            ctx.pad(cur)
            out = cur.download_padded()
            assert np.array_equal(out.buf_y, z[f"{name}_dbk_y"]) and np.array_equal(out.buf_u, z[f"{name}_dbk_u"]) and np.array_equal(out.buf_v, z[f"{name}_dbk_v"]), name
        for p in drefs + [cur]:
            p.free()


def test_golden_main_pipeline_gpu(ctx):
    """BASELINE config 3 in miniature through the C ABI on the GPU - xb200_recon_frame (all Main tools) -> xb200_deblock (ADDB, with
    the maps the reconstruction published) -> xb200_alf -> xb200_pad - against the recorded output of the unmodified reference"""
    from pathlib import Path
    from tests.test_golden import MAIN_CFGS
    z = np.load(Path(__file__).resolve().parent / "golden" / "main_frames.npz")
    w, h = 256, 136
    for name, kw in MAIN_CFGS:
        prm, cl, refs, alf, flags = synth.make_main_frame(w, h, **kw)
        drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
        cur = ctx.pic_alloc(w, h)
        ctx.set_chroma_qp_table(synth.chroma_qp_table(True))
        ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got = cur.download(maps=True)
        for pl, k in zip(got.planes(), "yuv"):
            assert np.array_equal(pl, z[f"{name}_rec_{k}"]), (name, k)
        assert np.array_equal(got.map_mv, z[f"{name}_map_mv"]) and np.array_equal(got.map_scu, z[f"{name}_map_scu"]), name
        ctx.deblock(prm, cur, drefs, drefs[::-1])
        ctx.alf(prm, cur, alf, flags)
        ctx.pad(cur)
        out = cur.download_padded()
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in drefs + [cur]:
            p.free()
        assert np.array_equal(out.buf_y, z[f"{name}_fin_y"]) and np.array_equal(out.buf_u, z[f"{name}_fin_u"]) and np.array_equal(out.buf_v, z[f"{name}_fin_v"]), name


def test_golden_tree_pipeline_gpu(ctx):
    """local dual tree nodes / constrained intra prediction through the C ABI - xb200_recon_frame -> xb200_deblock (maps and edge map as
    the reconstruction published them) -> xb200_pad - against the recorded output of the unmodified reference (tests/golden/tree_frames.npz)"""
    from pathlib import Path
    from tests.test_golden import TREE_CFGS, golden_tree_inputs
    z = np.load(Path(__file__).resolve().parent / "golden" / "tree_frames.npz")
    for name, kw, o in TREE_CFGS:
        w, h, prm, cl, refs = golden_tree_inputs(kw, o)
        drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
        cur = ctx.pic_alloc(w, h)
        ctx.set_chroma_qp_table(synth.chroma_qp_table(True))
        ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got = cur.download(maps=True)
        for pl, k in zip(got.planes(), "yuv"):
            assert np.array_equal(pl, z[f"{name}_rec_{k}"]), (name, k)
        assert np.array_equal(got.map_scu, z[f"{name}_map_scu"]), name
        ctx.deblock(prm, cur, drefs, drefs[::-1])
        ctx.pad(cur)
        out = cur.download_padded()
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in drefs + [cur]:
            p.free()
        assert np.array_equal(out.buf_y, z[f"{name}_fin_y"]) and np.array_equal(out.buf_u, z[f"{name}_fin_u"]) and np.array_equal(out.buf_v, z[f"{name}_fin_v"]), name


@pytest.mark.parametrize("variant,lg", [("B", 6), ("C", 6), ("C", 7), ("C", 5)])
def test_band_mode_single_gpu(ctx, oracle, variant, lg):
    """band mode of xb200_recon_frame (BASELINE config 4): the picture reconstructed as three separate CTU-row bands into a second
    device picture, moved over by xb200_band_pack / xb200_band_unpack, equals the whole-picture call (planes and maps)"""
    import torch
    from xevd_b200 import dist as xdist
    w, h, bd = 256, 328, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=91, n_refs=2, coded_frac=0.8, log2_ctu=lg)
    refs = synth.make_refs(w, h, bd, 2, seed=92)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    part, whole = ctx.pic_alloc(w, h), ctx.pic_alloc(w, h)
    ctu = 1 << lg
    for r0, k in xdist.band_partition(h, lg, 3):
        prm.ctu_row0, prm.ctu_rows = r0, k
        ctx.recon_frame(prm, part, drefs, drefs[::-1], cl.band(r0, k))
        y0, rows = r0 * ctu, min(k * ctu, h - r0 * ctu)
        buf = torch.empty(ctx.band_bytes(part, rows), dtype=torch.uint8, device="cuda")
        ctx.band_pack(part, y0, rows, buf.data_ptr())
        ctx.band_unpack(whole, y0, rows, buf.data_ptr())
        ctx.sync()
    prm.ctu_row0 = prm.ctu_rows = 0
    got = whole.download(maps=True)
    edge_got = whole.download_edge_map()
    ctx.recon_frame(prm, part, drefs, drefs[::-1], cl)
    edge_want = part.download_edge_map()
    for p in drefs + [part, whole]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(edge_got, edge_want)


@pytest.mark.parametrize("w,h", [(3840, 2160), (7680, 4320)])
def test_full_size_identity_property(ctx, w, h):
    """BASELINE full sizes (4K, 8K) through a size-independent property: zero motion + no coefficients must return the reference
    picture exactly (every CTU, every tile, the TMA windows and the store path are exercised; no oracle run needed)"""
    rng = np.random.default_rng(1)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=10, variant="A", seed=2, n_refs=1, coded_frac=0.0)
    cl.cus["mv"] = 0
    ref = HostPicture(w, h, 0)
    ref.y[...] = rng.integers(0, 1024, (h, w), dtype=np.int16)
    ref.u[...] = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.int16)
    ref.v[...] = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.int16)
    dref = ctx.pic_alloc(w, h).upload(ref)
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, [dref], [], cl)
    got = cur.download()
    dref.free(); cur.free()
    for a, b, n in zip(got.planes(), ref.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"


def test_full_size_integer_shift_property(ctx):
    """4K: whole-sample motion (8, -4) with no coefficients = the reference picture displaced (interior samples)"""
    w, h = 3840, 2160
    rng = np.random.default_rng(3)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=10, variant="B", seed=4, n_refs=1, coded_frac=0.0, bi_frac=0.0)
    cl.cus["mv"] = 0
    cl.cus["mv"][:, 0, 0] = 8 * 4
    cl.cus["mv"][:, 0, 1] = -4 * 4
    cl.cus["refi"][:, 0] = 0
    cl.cus["refi"][:, 1] = -1
    ref = HostPicture(w, h, 0)
    ref.y[...] = rng.integers(0, 1024, (h, w), dtype=np.int16)
    ref.u[...] = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.int16)
    ref.v[...] = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.int16)
    dref = ctx.pic_alloc(w, h).upload(ref)
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, [dref], [], cl)
    got = cur.download()
    dref.free(); cur.free()
    assert np.array_equal(got.y[4:, :w - 8], ref.y[:h - 4, 8:])
    assert np.array_equal(got.u[2:, :w // 2 - 4], ref.u[:h // 2 - 2, 4:]) and np.array_equal(got.v[2:, :w // 2 - 4], ref.v[:h // 2 - 2, 4:])


def test_4k_parity_vs_oracle(ctx, oracle):
    """the bench workload itself (4K config 2A and 2B) bit-exact against the oracle"""
    for variant in ("A", "B"):
        _run(ctx, oracle, 3840, 2160, 10, variant, seed=11, n_refs=2 if variant == "B" else 1)


def test_empty_and_ragged_inputs(ctx, oracle):
    """CTUs without any CU (ragged ctu_first) and a picture with no CU at all leave the picture untouched / succeed"""
    w, h, bd = 256, 136, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=5, n_refs=2)
    refs = synth.make_refs(w, h, bd, 2, seed=6)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    # drop the CUs of every second CTU
    keep = np.ones(cl.n_cu, bool)
    for k in range(0, cl.n_ctu, 2):
        keep[int(cl.ctu_first[k]):int(cl.ctu_first[k + 1])] = False
    sub = cl.band(0, (h + 63) // 64)
    counts = np.array([keep[int(cl.ctu_first[k]):int(cl.ctu_first[k + 1])].sum() for k in range(cl.n_ctu)])
    sub.cus = cl.cus[keep].copy()
    sub.ctu_first = np.concatenate(([0], np.cumsum(counts))).astype(np.uint32)
    sizes = np.array([__import__("xevd_b200.frame", fromlist=["cu_coef_count"]).cu_coef_count(c) for c in cl.cus])
    pieces = [cl.coef[int(c["coef_off"]):int(c["coef_off"]) + int(s)] for c, s, kf in zip(cl.cus, sizes, keep) if kf]
    sub.coef = np.concatenate(pieces) if pieces else np.zeros(0, np.int16)
    sub.cus["coef_off"] = np.concatenate(([0], np.cumsum(sizes[keep])))[:-1]
    sub.validate()
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], sub)
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], sub)
    got = cur.download()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    # no CU at all
    empty = cl.band(0, (h + 63) // 64)
    empty.cus = cl.cus[:0].copy(); empty.coef = np.zeros(0, np.int16); empty.ctu_first = np.zeros(cl.n_ctu + 1, np.uint32)
    cur2 = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur2, drefs, drefs[::-1], empty)
    assert not cur2.download().y.any()
    for p in drefs + [cur, cur2]:
        p.free()


@pytest.mark.parametrize("lg,iqt,seed", [(6, True, 31), (7, False, 32)])
def test_4k_main_full_pipeline(ctx, oracle, lg, iqt, seed):
    """BASELINE config 3 at its full size: a 3840x2160 10-bit Main-profile picture with every hot-path tool on (BTT + SUCO partition,
    1/16-pel taps, IQT, ATS, EIPD, IBC, HTDF, DMVR, affine) through xb200_recon_frame -> xb200_deblock (ADDB) -> xb200_alf -> xb200_pad,
    planes (with borders) and per-SCU maps bit-exact against the oracle"""
    w, h = 3840, 2160
    prm, cl, refs, alf, flags = synth.make_main_frame(w, h, bit_depth=10, seed=seed, log2_ctu=lg, iqt=iqt)
    cl.validate()
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.set_chroma_qp_table(synth.chroma_qp_table(True))
    try:
        ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got = cur.download(maps=True)
        for a, b, n in zip(got.planes(), want.planes(), "YUV"):
            assert np.array_equal(a, b), f"recon plane {n}: {int((a != b).sum())} samples differ"
        assert np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi) and np.array_equal(got.map_scu, want.map_scu)
        assert np.array_equal(got.map_unrefined_mv, want.map_unrefined_mv)
        oracle.deblock_frame(prm, want, cl, synth.chroma_qp_table(True), True, ((0, 1), (1, 0)))
        oracle.alf_frame(prm, want, alf, flags)
        oracle.pad(want)
        ctx.deblock(prm, cur, drefs, drefs[::-1])
        ctx.alf(prm, cur, alf, flags)
        ctx.pad(cur)
        out = cur.download_padded()
        assert np.array_equal(out.buf_y, want.buf_y) and np.array_equal(out.buf_u, want.buf_u) and np.array_equal(out.buf_v, want.buf_v)
    finally:
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in drefs + [cur]:
            p.free()


def test_8k_parity_vs_oracle(ctx, oracle):
    """BASELINE config 4's picture size on one GPU: 7680x4320 10-bit config 2A bit-exact against the oracle (planes and maps)"""
    _run(ctx, oracle, 7680, 4320, 10, "A", seed=12, n_refs=1)


def test_host_entry_refuses_malformed_work_lists(ctx):
    """xb200_recon_frame checks everything the kernels turn into an address: a malformed list fails the call with
    XB200_ERR_INVALID_ARGUMENT and a message naming the CU, instead of writing outside the picture or overrunning shared memory"""
    import copy
    from xevd_b200 import abi
    from xevd_b200.device import XevdB200Error
    w, h, bd = 256, 128, 10
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=5, n_refs=1, bi_frac=0.0)
    refs = synth.make_refs(w, h, bd, 1, seed=6)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, [], cl)            # the untouched list is accepted

    def refused(mutate, needle):
        bad = copy.deepcopy(cl)
        bad.cus = bad.cus.copy(); bad.ctu_first = bad.ctu_first.copy()
        mutate(bad)
        with pytest.raises(XevdB200Error) as e:
            ctx.recon_frame(prm, cur, drefs, [], bad)
        assert e.value.code == abi.XB200_ERR_INVALID_ARGUMENT, e.value
        assert needle in str(e.value), str(e.value)

    def m_first(b): b.ctu_first[-1] += 1
    def m_outside(b): b.cus["x"][0] = 64                 # CU 0 is listed under CTU 0
    def m_size(b): b.cus["log2w"][0] = 9
    def m_coef(b): b.cus["coef_off"][3] += 8
    def m_end(b): b.coef = b.coef[:-64]
    def m_ext(b):
        b.cus["mode"][0] = abi.MODE_AFFINE
        b.cus["mv"][0, 1] = np.frombuffer(np.uint32(77).tobytes(), np.int16)
    def m_ibc(b):
        b.cus["mode"][0] = abi.MODE_IBC
        b.cus["refi"][0] = -1
        b.cus["mv"][0, 0] = (-300, 0)
    refused(m_first, "ctu_first")
    refused(m_outside, "not inside the CTU")
    refused(m_size, "size outside")
    refused(m_coef, "coef_off")
    refused(m_end, "past the end")
    refused(m_ext, "extension record")
    refused(m_ibc, "block vector")
    ctx.recon_frame(prm, cur, drefs, [], cl)            # and the context is still usable
    ctx.sync()
    for p in drefs + [cur]:
        p.free()


@pytest.mark.parametrize("kind", ["inter-B", "main-all", "empty-coef"])
def test_sparse_coefficient_stream(ctx, oracle, kind):
    """xb200_recon_frame_sparse: (position, level) entries per 4096-coefficient chunk, expanded on the device - same pictures and maps as
    the dense call, against the oracle (what a caller behind PCIe sends: the dense stream is ~90 % zeros)"""
    from xevd_b200.frame import sparse_coef
    w, h, bd = 320, 192, 10
    if kind == "main-all":
        prm, cl, refs, _, _ = synth.make_main_frame(w, h, bit_depth=bd, seed=21)
        r0, r1 = refs, refs[::-1]
    else:
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=8, n_refs=2, coded_frac=0.0 if kind == "empty-coef" else 0.7)
        refs = synth.make_refs(w, h, bd, 2, seed=9)
        r0, r1 = refs, refs[::-1]
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), r0, r1, cl)
    entries, chunk_first = sparse_coef(cl.coef)
    assert entries.size <= cl.coef.size and chunk_first[-1] == entries.size
    d0 = [ctx.pic_alloc(w, h).upload(r) for r in r0]
    d1 = d0[::-1]
    cur = ctx.pic_alloc(w, h)
    for _ in range(2):                                   # twice: the staging ring reuses its slots
        ctx.recon_frame_sparse(prm, cur, d0, d1, cl, (entries, chunk_first))
    got = cur.download(maps=True)
    for p in d0 + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv)
