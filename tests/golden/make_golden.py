#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref (needs /root/reference once).

The reference ships no test vectors, and there are no EVC bitstreams on the build box, so the golden
vectors are outputs of the reference's own dispatched (AVX2) code on small seeded inputs, pushed through the
harness oracle/ref_harness.c.  Inputs are regenerated from the recorded seeds by xevd_b200.synth; outputs are stored.
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.pyoracle import Reference  # noqa: E402
from xevd_b200 import synth  # noqa: E402
from xevd_b200.frame import HostPicture  # noqa: E402

OUT = Path(__file__).resolve().parent


def golden_itdq(ref):
    out = {}
    for iqt in (0, 1):
        for bd in (8, 10):
            qp = 32 + 6 * (bd - 8)
            for lw in range(1, 7):
                for lh in range(1, 7):
                    rng = np.random.default_rng(1000 * iqt + 100 * bd + 10 * lw + lh)
                    res = rng.laplace(0, 40.0 * (1 << (bd - 8)), (1 << lh, 1 << lw))
                    lev = synth.quantised_dct(res, qp, bool(iqt))
                    out[f"in_{iqt}_{bd}_{lw}_{lh}"] = lev
                    out[f"out_{iqt}_{bd}_{lw}_{lh}"] = ref.itdq_block(lev, qp, bd, iqt)
    np.savez_compressed(OUT / "itdq_blocks.npz", **out)


def golden_mc(ref):
    out = {}
    rng = np.random.default_rng(77)
    for bd in (8, 10):
        plane = rng.integers(0, 1 << bd, (96, 112), dtype=np.int16)
        out[f"plane_{bd}"] = plane
        cases = []
        for t in range(48):
            main = t & 1
            w, h = int(rng.choice([4, 8, 16, 32])), int(rng.choice([4, 8, 16, 32]))
            step = 1 if main else 4
            chroma = (t >> 1) & 1
            n = 32 if chroma else 16
            fx, fy = int(rng.integers(0, n // step)) * step, int(rng.integers(0, n // step)) * step
            sh = 5 if chroma else 4
            gx, gy = (int(rng.integers(0, 40)) << sh) + fx, (int(rng.integers(0, 30)) << sh) + fy
            ox, oy = (fx, fy) if t % 5 else (int(rng.integers(0, n)), int(rng.integers(0, n)))
            cases.append((w, h, gx, gy, ox, oy, chroma, main))
            out[f"mc_{bd}_{t}"] = ref.mc(plane, (24, 20), (gx, gy), (ox, oy), w, h, bd, bool(chroma), bool(main))
        out[f"cases_{bd}"] = np.array(cases, np.int32)
    np.savez_compressed(OUT / "mc_blocks.npz", **out)


def golden_frames(ref):
    out = {}
    w, h = 128, 72
    cfgs = [("inter_A_10", dict(variant="A", bit_depth=10), 0.0), ("inter_B_8", dict(variant="B", bit_depth=8), 0.0),
            ("inter_B_10_iqt", dict(variant="B", bit_depth=10, iqt=True), 0.0), ("mixed_B_10", dict(variant="B", bit_depth=10), 0.4),
            ("intra_A8_10", dict(variant="A", bit_depth=10, log2_cu=3), 1.0)]
    for name, kw, intra in cfgs:
        bd = kw["bit_depth"]
        prm, cl = synth.make_inter_frame(w, h, seed=5, n_refs=2, coded_frac=0.7, **kw)
        if intra > 0:
            synth.add_intra_cus(cl, np.random.default_rng(6), intra)
        refs = synth.make_refs(w, h, bd, 2, seed=9)
        pic = ref.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        out[f"{name}_y"], out[f"{name}_u"], out[f"{name}_v"] = pic.y.copy(), pic.u.copy(), pic.v.copy()
        # deblocking of the same picture (Baseline filter) with per-CU QPs, then border padding
        if not kw.get("iqt"):
            work = pic.copy()
            for pl in work.planes():
                pl[...] = (pl.astype(np.int32) // 8 + (1 << (bd - 1))).astype(np.int16)
            maps = HostPicture(w, h)
            synth.randomize_deblock_maps(work, cl, np.random.default_rng(8))
            ref.deblock_frame(prm, work, cl, synth.chroma_qp_table(False))
            ref.pad(work)
            out[f"{name}_dbk_y"], out[f"{name}_dbk_u"], out[f"{name}_dbk_v"] = work.buf_y.copy(), work.buf_u.copy(), work.buf_v.copy()
    np.savez_compressed(OUT / "frames.npz", **out)


MAIN_CFGS = [("main_ctu64_iqt_10", dict(bit_depth=10, seed=3, log2_ctu=6, iqt=True)), ("main_ctu128_10", dict(bit_depth=10, seed=4, log2_ctu=7, iqt=False)),
             ("main_ctu32_iqt_8", dict(bit_depth=8, seed=5, log2_ctu=5, iqt=True))]


def golden_main(ref):
    """BASELINE config 3 in miniature: every Main-profile hot-path tool on, through the reference's own per-CU calls, then the
    reference's ADDB deblocking, ALF and border padding.  Per-SCU maps between the stages are the oracle's (the harness has no
    slice-level state); the oracle's maps are pinned separately (tests/test_oracle_vs_ref.py)."""
    from oracle.pyoracle import Oracle
    orc = Oracle()
    out = {}
    w, h = 256, 136
    for name, kw in MAIN_CFGS:
        prm, cl, refs, alf, flags = synth.make_main_frame(w, h, **kw)
        pic = ref.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        o = orc.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        out[f"{name}_rec_y"], out[f"{name}_rec_u"], out[f"{name}_rec_v"] = pic.y.copy(), pic.u.copy(), pic.v.copy()
        out[f"{name}_map_mv"], out[f"{name}_map_scu"] = o.map_mv.copy(), o.map_scu.copy()
        for k in ("map_mv", "map_refi", "map_scu", "map_unrefined_mv"):
            getattr(pic, k)[...] = getattr(o, k)
        ref.deblock_frame(prm, pic, cl, synth.chroma_qp_table(True), True, ((0, 1), (1, 0)))
        ref.alf_frame(prm, pic, alf, flags)
        ref.pad(pic)
        out[f"{name}_fin_y"], out[f"{name}_fin_u"], out[f"{name}_fin_v"] = pic.buf_y.copy(), pic.buf_u.copy(), pic.buf_v.copy()
    np.savez_compressed(OUT / "main_frames.npz", **out)


def golden_tree(ref):
    """local dual tree nodes and constrained intra prediction: the reference's per-CU calls with its own tree_cons / constrained_intra_flag
    arguments, then its deblocking (Baseline filter or ADDB, per-CU tree_cons) and border padding.  Maps between the stages are the
    oracle's, as in golden_main."""
    from oracle.pyoracle import Oracle
    from tests.test_golden import TREE_CFGS, golden_tree_inputs
    orc = Oracle()
    out = {}
    for name, kw, o in TREE_CFGS:
        w, h, prm, cl, refs = golden_tree_inputs(kw, o)
        pic = ref.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        m = orc.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        out[f"{name}_rec_y"], out[f"{name}_rec_u"], out[f"{name}_rec_v"] = pic.y.copy(), pic.u.copy(), pic.v.copy()
        out[f"{name}_map_scu"] = m.map_scu.copy()
        for k in ("map_mv", "map_refi", "map_scu", "map_unrefined_mv"):
            getattr(pic, k)[...] = getattr(m, k)
        ref.deblock_frame(prm, pic, cl, synth.chroma_qp_table(True), bool(o["addb"]), ((0, 1), (1, 0)))
        ref.pad(pic)
        out[f"{name}_fin_y"], out[f"{name}_fin_u"], out[f"{name}_fin_v"] = pic.buf_y.copy(), pic.buf_u.copy(), pic.buf_v.copy()
    np.savez_compressed(OUT / "tree_frames.npz", **out)


if __name__ == "__main__":
    r = Reference(2)
    only = sys.argv[1:] or ["itdq", "mc", "frames", "main", "tree"]
    for k, fn in (("itdq", golden_itdq), ("mc", golden_mc), ("frames", golden_frames), ("main", golden_main), ("tree", golden_tree)):
        if k in only:
            fn(r)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")
