"""The CPU oracle against the committed golden vectors (tests/golden/*.npz), which are outputs of the unmodified reference
(dispatched AVX2 code) recorded by tests/golden/make_golden.py.  Runs anywhere: needs neither /root/reference nor a GPU."""
from pathlib import Path

import numpy as np
import pytest

from xevd_b200 import synth
from xevd_b200.frame import HostPicture

G = Path(__file__).resolve().parent / "golden"


def test_golden_itdq(oracle):
    z = np.load(G / "itdq_blocks.npz")
    n = 0
    for key in z.files:
        if not key.startswith("in_"):
            continue
        _, iqt, bd, lw, lh = key.split("_")
        out = oracle.itdq_block(z[key], 32 + 6 * (int(bd) - 8), int(bd), int(iqt))
        assert np.array_equal(out, z["out" + key[2:]]), key
        n += 1
    assert n == 2 * 2 * 36


def test_golden_mc(oracle):
    z = np.load(G / "mc_blocks.npz")
    for bd in (8, 10):
        plane = z[f"plane_{bd}"]
        for t, (w, h, gx, gy, ox, oy, chroma, main) in enumerate(z[f"cases_{bd}"]):
            got = oracle.mc(plane, (24, 20), (int(gx), int(gy)), (int(ox), int(oy)), int(w), int(h), bd, bool(chroma), bool(main))
            assert np.array_equal(got, z[f"mc_{bd}_{t}"]), (bd, t)


FRAME_CFGS = [("inter_A_10", dict(variant="A", bit_depth=10), 0.0), ("inter_B_8", dict(variant="B", bit_depth=8), 0.0),
              ("inter_B_10_iqt", dict(variant="B", bit_depth=10, iqt=True), 0.0), ("mixed_B_10", dict(variant="B", bit_depth=10), 0.4),
              ("intra_A8_10", dict(variant="A", bit_depth=10, log2_cu=3), 1.0)]


def golden_frame_inputs(name, kw, intra):
    """the seeded inputs make_golden.py used"""
    w, h = 128, 72
    prm, cl = synth.make_inter_frame(w, h, seed=5, n_refs=2, coded_frac=0.7, **kw)
    if intra > 0:
        synth.add_intra_cus(cl, np.random.default_rng(6), intra)
    refs = synth.make_refs(w, h, kw["bit_depth"], 2, seed=9)
    return w, h, prm, cl, refs


@pytest.mark.parametrize("name,kw,intra", FRAME_CFGS)
def test_golden_frames(oracle, name, kw, intra):
    z = np.load(G / "frames.npz")
    w, h, prm, cl, refs = golden_frame_inputs(name, kw, intra)
    bd = kw["bit_depth"]
    pic = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl, k in zip(pic.planes(), "yuv"):
        assert np.array_equal(pl, z[f"{name}_{k}"]), (name, k)
    if not kw.get("iqt"):
        for pl in pic.planes():
            pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
        synth.randomize_deblock_maps(pic, cl, np.random.default_rng(8))
        oracle.deblock_frame(prm, pic, cl, synth.chroma_qp_table(False))
        oracle.pad(pic)
        assert np.array_equal(pic.buf_y, z[f"{name}_dbk_y"]) and np.array_equal(pic.buf_u, z[f"{name}_dbk_u"]) and np.array_equal(pic.buf_v, z[f"{name}_dbk_v"])


MAIN_CFGS = [("main_ctu64_iqt_10", dict(bit_depth=10, seed=3, log2_ctu=6, iqt=True)), ("main_ctu128_10", dict(bit_depth=10, seed=4, log2_ctu=7, iqt=False)),
             ("main_ctu32_iqt_8", dict(bit_depth=8, seed=5, log2_ctu=5, iqt=True))]


@pytest.mark.parametrize("name,kw", MAIN_CFGS)
def test_golden_main_pipeline(oracle, name, kw):
    """BASELINE config 3 in miniature (all Main tools, then ADDB deblocking + ALF + padding) against the reference's recorded output"""
    z = np.load(G / "main_frames.npz")
    w, h = 256, 136
    prm, cl, refs, alf, flags = synth.make_main_frame(w, h, **kw)
    pic = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl, k in zip(pic.planes(), "yuv"):
        assert np.array_equal(pl, z[f"{name}_rec_{k}"]), (name, k)
    assert np.array_equal(pic.map_mv, z[f"{name}_map_mv"]) and np.array_equal(pic.map_scu, z[f"{name}_map_scu"])
    oracle.deblock_frame(prm, pic, cl, synth.chroma_qp_table(True), True, ((0, 1), (1, 0)))
    oracle.alf_frame(prm, pic, alf, flags)
    oracle.pad(pic)
    assert np.array_equal(pic.buf_y, z[f"{name}_fin_y"]) and np.array_equal(pic.buf_u, z[f"{name}_fin_u"]) and np.array_equal(pic.buf_v, z[f"{name}_fin_v"])


TREE_CFGS = [("dual_eipd_htdf_10", dict(bit_depth=10, seed=11, log2_ctu=6), dict(eipd=1, htdf=1, constrained=0, intra=1.0, addb=0)),
             ("dual_base_8", dict(bit_depth=8, seed=12, log2_ctu=5), dict(eipd=0, htdf=0, constrained=0, intra=0.6, addb=0)),
             ("dual_constrained_iqt_10", dict(bit_depth=10, seed=13, log2_ctu=7, iqt=True), dict(eipd=1, htdf=1, constrained=1, intra=0.5, addb=1)),
             ("constrained_12", dict(bit_depth=12, seed=14, log2_ctu=6), dict(eipd=1, htdf=1, constrained=1, intra=0.4, addb=1, dual=0))]


def golden_tree_inputs(kw, o):
    """the seeded inputs make_golden.py used for tree_frames.npz: BTT partitions with local dual tree nodes (luma-only leaves + one
    chroma-only CU, src_main/xevdm.c:1828-1846) and / or pps.constrained_intra_pred_flag in mixed intra / inter pictures"""
    w, h = 256, 136
    kw = dict(kw)
    seed = kw.pop("seed")
    prm, cl = synth.make_inter_frame(w, h, variant="C", seed=seed, n_refs=2, coded_frac=0.7, **kw)
    prm.tool_eipd, prm.tool_htdf, prm.slice_qp, prm.constrained_intra_pred, prm.tool_addb = o["eipd"], o["htdf"], 37, o["constrained"], o["addb"]
    prm.qp_u_offset, prm.qp_v_offset = 1, -2
    if o.get("dual", 1):
        synth.split_local_dual_tree(cl, np.random.default_rng(seed + 1), 0.6)
    synth.add_intra_cus(cl, np.random.default_rng(seed + 2), o["intra"], eipd=bool(o["eipd"]), constrained=bool(o["constrained"]))
    synth.derive_avail_cu(cl)
    cl.validate()
    return w, h, prm, cl, synth.make_refs(w, h, kw["bit_depth"], 2, seed=seed + 3)


@pytest.mark.parametrize("name,kw,o", TREE_CFGS)
def test_golden_tree_pipeline(oracle, name, kw, o):
    """recon -> deblock -> pad of pictures with local dual tree nodes / constrained intra prediction against the recorded output of the
    unmodified reference (its own tree_cons and constrained_intra_flag arguments)"""
    z = np.load(G / "tree_frames.npz")
    w, h, prm, cl, refs = golden_tree_inputs(kw, o)
    pic = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    for pl, k in zip(pic.planes(), "yuv"):
        assert np.array_equal(pl, z[f"{name}_rec_{k}"]), (name, k)
    assert np.array_equal(pic.map_scu, z[f"{name}_map_scu"])
    oracle.deblock_frame(prm, pic, cl, synth.chroma_qp_table(True), bool(o["addb"]), ((0, 1), (1, 0)))
    oracle.pad(pic)
    assert np.array_equal(pic.buf_y, z[f"{name}_fin_y"]) and np.array_equal(pic.buf_u, z[f"{name}_fin_u"]) and np.array_equal(pic.buf_v, z[f"{name}_fin_v"])
