#!/usr/bin/env python3
"""Per-pass timing of the whole hot path on one GPU (BASELINE configs 2, 3 and 5): every kernel family on 4K 10-bit pictures, CUDA
events per call on the launching stream, NPIC distinct pictures in rotation (> L2), algorithmic bytes per SURVEY 8(d).
    python tools/bench_passes.py [--w 3840 --h 2160] [--reps 5] > profiles/rN/passes.json
One JSON object per line: {"pass", "us", "alg_bytes", "gbs", "frac_hbm"}."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    from xevd_b200 import synth
    from xevd_b200.device import Context
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--npic", type=int, default=6)
    args = ap.parse_args()
    w, h, bd = args.w, args.h, 10
    peak = 6650.0
    p = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(0)
    ctx.set_stream(stream.cuda_stream)
    samples = w * h * 3 // 2
    nscu = (w // 4) * (h // 4)

    def timed(name, fn, alg_bytes, n=args.npic):
        """fn(i) launches the pass on picture slot i; every slot is touched once per repetition"""
        for i in range(n):
            fn(i)
        ctx.sync()
        ts = []
        for _ in range(args.reps):
            for i in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(i); e1.record(stream)
                ts.append((e0, e1))
        ctx.sync()
        us = 1e3 * float(np.median([a.elapsed_time(b) for a, b in ts]))
        gbs = alg_bytes / (us * 1e-6) / 1e9
        print(json.dumps({"pass": name, "picture": f"{w}x{h}", "us": round(us, 1), "alg_bytes": int(alg_bytes), "gbs": round(gbs, 1),
                          "frac_hbm": round(gbs / peak, 4), "peak_gbs": peak}), flush=True)

    def upload_work(cl):
        return dict(cl=cl, cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev), first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev),
                    ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev), coef=torch.from_numpy(cl.coef.copy()).to(dev),
                    max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max()))

    def recon(prm, cur, refs, refs1, wk, has_intra):
        cl = wk["cl"]
        ctx.recon_frame_dev(prm, cur, refs, refs1, wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(), cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext),
                            wk["coef"].data_ptr(), cl.coef.size, has_intra=has_intra, max_cu_per_ctu=wk["max_cu"])

    def alg_recon(cl):
        c = cl.cus
        smp = (1 << (c["log2w"].astype(np.int64) + c["log2h"].astype(np.int64))) * 3 // 2
        ndir = np.where(c["mode"] == 0, 0, (c["refi"][:, 0] >= 0).astype(np.int64) + (c["refi"][:, 1] >= 0).astype(np.int64))
        ndir = np.where(c["mode"] == 4, 1, ndir)
        return int((smp * ndir).sum()) * 2 + cl.coef.size * 2 + int(smp.sum()) * 2 + c.nbytes + cl.ctu_first.nbytes

    host_refs = synth.make_refs(w, h, bd, 2, seed=7)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in host_refs]
    curs = [ctx.pic_alloc(w, h) for _ in range(args.npic)]

    # config 2A / 2B: Baseline inter pictures through the throughput kernel
    for variant in ("A", "B"):
        t0 = time.time()
        works = []
        for i in range(2):
            prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=1 + i, n_refs=1 if variant == "A" else 2)
            works.append((prm, upload_work(cl)))
        r1 = [] if variant == "A" else drefs[::-1]
        r0 = drefs[:1] if variant == "A" else drefs
        timed(f"recon_inter_v2 (config 2{variant})", lambda i: recon(works[i % 2][0], curs[i], r0, r1, works[i % 2][1], False), alg_recon(works[0][1]["cl"]))
    # Main-profile inter pictures: IQT + 1/16-pel tables go through the throughput kernel as well; with ATS / DMVR / affine enabled the CTUs
    # that hold such CUs go through the generic kernel (per-CU dispatch, two launches)
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=1, n_refs=2, iqt=True, main_mv=True)
    wk = upload_work(cl)
    timed("recon_inter_v2 (IQT, 1/16-pel, quadtree)", lambda i: recon(prm, curs[i], drefs, drefs[::-1], wk, False), alg_recon(cl))
    prm_m, cl_m, refs_m = synth.make_dmvr_case(w, h, bit_depth=bd, variant="C", seed=5, coded_frac=0.6, main_mv=True, ats_inter_frac=0.3, iqt=True)
    prm_m.tool_affine = 1
    synth.add_affine_cus(cl_m, np.random.default_rng(3), 0.3)
    dm = [ctx.pic_alloc(w, h).upload(r) for r in refs_m]
    wk_m = upload_work(cl_m)
    timed("recon Main picture: BTT, ATS, DMVR, affine (v2 + generic, per-CU dispatch)", lambda i: recon(prm_m, curs[i], dm, dm[::-1], wk_m, False), alg_recon(cl_m))
    # the same tools on a few per cent of the CUs: most CTUs stay with the throughput kernel
    prm_s, cl_s, refs_s = synth.make_dmvr_case(w, h, bit_depth=bd, variant="C", seed=6, flag_frac=0.03, coded_frac=0.6, main_mv=True, ats_inter_frac=0.01, iqt=True)
    prm_s.tool_affine = 1
    synth.add_affine_cus(cl_s, np.random.default_rng(4), 0.01)
    ds_ = [ctx.pic_alloc(w, h).upload(r) for r in refs_s]
    wk_s = upload_work(cl_s)
    timed("recon Main picture: BTT, 1 % ATS, 1 % affine, 3 % DMVR flags (v2 + generic, per-CU dispatch)", lambda i: recon(prm_s, curs[i], ds_, ds_[::-1], wk_s, False), alg_recon(cl_s))
    # the same three pictures through the generic kernel alone
    import os
    os.environ["XB200_FORCE_GENERIC"] = "1"
    ctx_g = Context(0)
    del os.environ["XB200_FORCE_GENERIC"]
    ctx_g.set_stream(stream.cuda_stream)

    def recon_g(prm, cur, refs, refs1, wk):
        cl = wk["cl"]
        ctx_g.recon_frame_dev(prm, cur, refs, refs1, wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(), cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext),
                              wk["coef"].data_ptr(), cl.coef.size, has_intra=False, max_cu_per_ctu=wk["max_cu"])
    timed("recon_inter generic alone (IQT, 1/16-pel, quadtree)", lambda i: recon_g(prm, curs[i], drefs, drefs[::-1], wk), alg_recon(cl))
    timed("recon_inter generic alone (Main: BTT, ATS, DMVR, affine)", lambda i: recon_g(prm_m, curs[i], dm, dm[::-1], wk_m), alg_recon(cl_m))
    timed("recon_inter generic alone (Main: BTT, 1 % ATS, 1 % affine, 3 % DMVR flags)", lambda i: recon_g(prm_s, curs[i], ds_, ds_[::-1], wk_s), alg_recon(cl_s))
    # I picture: Baseline modes and EIPD + HTDF, through the wavefront kernel
    for eipd in (0, 1):
        prm_i, cl_i = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=9, n_refs=1, coded_frac=0.7, iqt=bool(eipd))
        prm_i.tool_eipd = eipd
        prm_i.tool_htdf = eipd
        prm_i.slice_qp = 34
        synth.add_intra_cus(cl_i, np.random.default_rng(2), 1.0, eipd=bool(eipd))
        synth.derive_avail_cu(cl_i)
        wk_i = upload_work(cl_i)
        timed("recon_intra wavefront (I picture, " + ("EIPD + HTDF" if eipd else "Baseline modes") + ")",
              lambda i: recon(prm_i, curs[i], drefs[:1], [], wk_i, True), alg_recon(cl_i))
    # P pictures with a share of intra CUs: throughput inter kernel + wavefront kernel over the CTUs that hold intra CUs
    for frac in (0.02, 0.1):
        prm_x, cl_x = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=21, n_refs=2, coded_frac=0.7)
        synth.add_intra_cus(cl_x, np.random.default_rng(6), frac)
        synth.derive_avail_cu(cl_x)
        wk_x = upload_work(cl_x)
        timed(f"recon P picture, {int(frac * 100)} % intra CUs (inter v2 + wavefront)", lambda i: recon(prm_x, curs[i], drefs, drefs[::-1], wk_x, True), alg_recon(cl_x))
    # Main P picture with HTDF: every CU with a luma residual goes through the in-order filter of the wavefront kernel; uncoded CUs break
    # the dependency chains
    for cf in (0.15, 0.6):
        prm_h, cl_h = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=23, n_refs=2, coded_frac=cf, iqt=True, main_mv=True)
        prm_h.tool_eipd = prm_h.tool_htdf = 1
        prm_h.slice_qp = 34
        synth.add_intra_cus(cl_h, np.random.default_rng(7), 0.02, eipd=True)
        synth.derive_avail_cu(cl_h)
        wk_h = upload_work(cl_h)
        timed(f"recon Main P picture with HTDF, {int(cf * 100)} % of the CUs coded, 2 % intra (v2 + wavefront)", lambda i: recon(prm_h, curs[i], drefs, drefs[::-1], wk_h, True),
              alg_recon(cl_h))
    # picture-wide passes on reconstructed pictures (maps left by the Main inter reconstruction above)
    for i in range(args.npic):
        recon(prm_m, curs[i], dm, dm[::-1], wk_m, False)
    ctx.set_chroma_qp_table(synth.chroma_qp_table(True))
    dbk_bytes = samples * 4 + nscu * 15
    prm_b = type(prm_m).from_buffer_copy(prm_m)
    prm_b.tool_addb = 0
    timed("deblock Baseline filter (2 launches)", lambda i: ctx.deblock(prm_b, curs[i], dm, dm[::-1]), dbk_bytes)
    prm_a = type(prm_m).from_buffer_copy(prm_m)
    prm_a.tool_addb = 1
    timed("deblock ADDB (2 launches)", lambda i: ctx.deblock(prm_a, curs[i], dm, dm[::-1]), dbk_bytes)
    alf = synth.make_alf_params(np.random.default_rng(4))
    timed("ALF (copy + filter)", lambda i: ctx.alf(prm_m, curs[i], alf, None), samples * 4)
    timed("pad (xevd_picbuf_expand)", lambda i: ctx.pad(curs[i]), 2 * (2 * 144 * (w + 288) + 2 * 144 * h) + 4 * (2 * 72 * (w // 2 + 144) + 2 * 72 * (h // 2)))
    # config 5: leaf micro-benchmarks
    n_s = 1 << 22
    for lg in (2, 3, 4, 5, 6):
        nb = n_s >> (2 * lg)
        rng = np.random.default_rng(lg)
        lev = synth.quantised_dct(rng.laplace(0, 8.0, (nb, 1 << lg, 1 << lg)), 44)
        d_in = torch.from_numpy(lev.reshape(-1).copy()).to(dev)
        d_out = torch.empty_like(d_in)
        timed(f"itdq_blocks {1 << lg}x{1 << lg} (config 5)", lambda i: ctx.itdq_blocks_dev(d_in.data_ptr(), d_out.data_ptr(), nb, lg, lg, 44, bd, False), 4 * n_s, n=1)


if __name__ == "__main__":
    main()
