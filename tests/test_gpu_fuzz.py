"""Randomised whole-pipeline parity on the GPU: picture size (not CTU-aligned), bit depth, CTU size, partition, tool mix and densities
are drawn per seed; xb200_recon_frame -> xb200_deblock -> xb200_alf -> xb200_pad through the C ABI must equal the oracle's pipeline
(which is pinned to the unmodified reference by tests/test_oracle_vs_ref.py and the golden vectors) bit for bit, maps included."""
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


@pytest.fixture(scope="module")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


def draw_case(seed, tree=False):
    """tree: additionally local dual tree nodes (luma-only leaves + chroma-only CU) and / or pps.constrained_intra_pred_flag, drawn from a
    second generator so that the cases of the plain sweep stay what they were"""
    rng = np.random.default_rng(4000 + seed)
    rng2 = np.random.default_rng(9000 + seed)
    w, h = 8 * int(rng.integers(20, 72)), 8 * int(rng.integers(12, 44))
    bd = int(rng.choice([8, 10, 10, 12]))
    main = bool(rng.random() < 0.7) or tree           # the tree sweep is Main-profile only (BTT); the draw is kept so that the streams stay aligned
    lg = int(rng.choice([5, 6, 6, 6, 7])) if main else 6
    coded = float(rng.choice([0.15, 0.5, 0.9]))
    if not main:
        variant = str(rng.choice(["A", "B"]))
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=seed, n_refs=2, coded_frac=coded, log2_cu=int(rng.choice([3, 4, 5])))
        refs = synth.make_refs(w, h, bd, 2, seed=seed + 1)
        if rng.random() < 0.6:
            synth.add_intra_cus(cl, rng, float(rng.choice([0.03, 0.2, 1.0])))
        cl.validate()
        return dict(w=w, h=h, prm=prm, cl=cl, refs=refs, alf=None, flags=None, main=False)
    iqt = bool(rng.integers(0, 2))
    ats = float(rng.choice([0.0, 0.02, 0.3]))
    if rng.random() < 0.5:
        prm, cl, refs = synth.make_dmvr_case(w, h, bit_depth=bd, variant="C", seed=seed, flag_frac=float(rng.choice([0.03, 0.8])), coded_frac=coded, main_mv=True,
                                             ats_inter_frac=ats, iqt=iqt, log2_ctu=lg)
    else:
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="C", seed=seed, n_refs=2, coded_frac=coded, main_mv=True, ats_inter_frac=ats, iqt=iqt,
                                         log2_ctu=lg, suco=bool(rng.integers(0, 2)))
        refs = synth.make_refs(w, h, bd, 2, seed=seed + 1)
    prm.tool_eipd = 1
    prm.tool_htdf = int(rng.integers(0, 2))
    prm.tool_ibc = int(rng.random() < 0.3)
    prm.tool_addb = int(rng.integers(0, 2))
    prm.tool_alf = 1
    prm.slice_qp = int(rng.integers(18, 50))
    prm.qp_u_offset, prm.qp_v_offset = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
    cl.cus["qp_map"] = rng.integers(20, 48, cl.n_cu)
    intra = float(rng.choice([0.0, 0.03, 0.25, 1.0]))
    constrained = False
    if tree:
        constrained = bool(rng2.random() < 0.5)
        prm.constrained_intra_pred = int(constrained)
        if rng2.random() < 0.8:
            synth.split_local_dual_tree(cl, rng2, float(rng2.choice([0.05, 0.5, 1.0])))
    if intra or prm.tool_ibc or tree:
        synth.add_intra_cus(cl, rng, max(intra, 0.05), eipd=True, ats_intra_frac=0.5 if ats else 0.0, ibc_frac=0.2 if prm.tool_ibc else 0.0, constrained=constrained)
    if rng.random() < 0.5:
        prm.tool_affine = 1
        synth.add_affine_cus(cl, rng, float(rng.choice([0.02, 0.3])))
    synth.derive_avail_cu(cl)
    cl.validate()
    ctu = 1 << lg
    n_ctu = ((w + ctu - 1) // ctu) * ((h + ctu - 1) // ctu)
    return dict(w=w, h=h, prm=prm, cl=cl, refs=refs, alf=synth.make_alf_params(rng, enable=tuple(int(x) for x in rng.integers(0, 2, 3))),
                flags=(rng.random(n_ctu) < 0.7).astype(np.uint8), main=True)


@pytest.mark.parametrize("seed", range(32))
def test_random_pipeline_tree(ctx, oracle, seed):
    """the same sweep over Main pictures with local dual tree nodes and constrained intra prediction mixed into the tool set"""
    run_case(ctx, oracle, draw_case(seed, tree=True), seed)


@pytest.mark.parametrize("seed", range(64))
def test_random_pipeline(ctx, oracle, seed):
    run_case(ctx, oracle, draw_case(seed), seed)


def run_case(ctx, oracle, k, seed):
    w, h, prm, cl, refs = k["w"], k["h"], k["prm"], k["cl"], k["refs"]
    tbl = synth.chroma_qp_table(k["main"])
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    try:
        ctx.set_chroma_qp_table(tbl)
        ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
        got = cur.download(maps=True)
        for a, b, n in zip(got.planes(), want.planes(), "YUV"):
            assert np.array_equal(a, b), f"seed {seed} recon plane {n}: {int((a != b).sum())} samples differ"
        assert np.array_equal(got.map_scu, want.map_scu) and np.array_equal(got.map_mv, want.map_mv) and np.array_equal(got.map_refi, want.map_refi)
        oracle.deblock_frame(prm, want, cl, tbl, bool(prm.tool_addb), ((0, 1), (1, 0)))
        ctx.deblock(prm, cur, drefs, drefs[::-1])
        if k["alf"] is not None:
            oracle.alf_frame(prm, want, k["alf"], k["flags"])
            ctx.alf(prm, cur, k["alf"], k["flags"])
        oracle.pad(want)
        ctx.pad(cur)
        out = cur.download_padded()
        assert np.array_equal(out.buf_y, want.buf_y) and np.array_equal(out.buf_u, want.buf_u) and np.array_equal(out.buf_v, want.buf_v), f"seed {seed}: final picture differs"
    finally:
        ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
        for p in drefs + [cur]:
            p.free()
