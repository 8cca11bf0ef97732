"""bench_extra -- the non-headline workloads of bench.py: every number DESIGN.md / profiles quote, measured in the driver's own run.

bench.py (rank 0, one GPU) calls run() after the headline measurement and puts the returned list under "extra" in its JSON line.  Every
entry: the per-picture device time of the whole call sequence of that workload (CUDA events on the launching stream, inputs resident,
NPIC distinct picture slots in rotation so that a repetition streams more than the 126 MB L2), its algorithmic bytes and
HBM-roofline fraction, and the SAME pass through the reference's own CPU code (oracle/_ref/libxevd_ref.so, dispatched AVX2) on one thread
and - for the workloads where the question is open - on all host cores (one single-threaded instance per core).

    4k-2B, 4k-iqt, 1080p-2A, 8k-2A      inter pictures through xb200_recon_frame_dev + xb200_pad (BASELINE config 2)
    4k-main-full                        config 3: Main picture with every tool -> recon + ADDB deblock + ALF + pad
    4k-P-intra10, 4k-I-baseline,        pictures that need the CTU wavefront kernel; I pictures also with several pictures in flight
    4k-I-eipd-htdf                        on separate streams (independent pictures: what GOP-level sharding gives inside one GPU)
    config5-itdq / config5-mc           leaf kernels: all 36 transform shapes, 8-tap luma / 4-tap chroma `nn` at 8..64 squared
    stream-*                            whole-decoder drop-in: libxevd_gpu.so vs libxevd_ref.so through xevd_create/xevd_decode/xevd_pull
"""
from __future__ import annotations

import ctypes as C
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def alg_recon(cl):
    """each datum once - reference read 2 B per sample and prediction direction (IBC: 1 direction), coefficients, write"""
    c = cl.cus
    smp = (1 << (c["log2w"].astype(np.int64) + c["log2h"].astype(np.int64))) * 3 // 2
    ndir = np.where(c["mode"] == 0, 0, (c["refi"][:, 0] >= 0).astype(np.int64) + (c["refi"][:, 1] >= 0).astype(np.int64))
    ndir = np.where(c["mode"] == 4, 1, ndir)
    return int((smp * ndir).sum()) * 2 + cl.coef.size * 2 + int(smp.sum()) * 2 + c.nbytes + cl.ctu_first.nbytes


def _cpu_worker(fn_name, payload, reps, q, barrier):
    fn = globals()[fn_name](payload)
    fn()
    if barrier is not None:
        barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    q.put(time.perf_counter() - t0)


def cpu_rate(fn_name, payload, cores, reps=1):
    """units/s of `reps` calls per process on `cores` forked single-threaded processes (cores == 1: in-process)"""
    if cores == 1:
        fn = globals()[fn_name](payload)
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return reps / (time.perf_counter() - t0)
    ctx = mp.get_context("fork")
    q, bar = ctx.Queue(), ctx.Barrier(cores)
    ps = [ctx.Process(target=_cpu_worker, args=(fn_name, payload, reps, q, bar)) for _ in range(cores)]
    for p in ps:
        p.start()
    ts = [q.get() for _ in ps]
    for p in ps:
        p.join()
    return cores * reps / max(ts)


# ---- CPU closures (built inside the worker process: the reference library keeps global state) ----------------------------------------
def cpu_picture(payload):
    from oracle.pyoracle import Reference
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    prm, cl, refs, refs1, post = payload
    be = Reference(2)
    cur = HostPicture(prm.w, prm.h, prm.poc)
    tbl = synth.chroma_qp_table(bool(prm.tool_iqt))

    def fn():
        be.recon_frame(prm, cur, refs, refs1, cl)
        if post is not None:
            alf, flags, ids = post
            be.deblock_frame(prm, cur, cl, tbl, bool(prm.tool_addb), ids)
            be.alf_frame(prm, cur, alf, flags)
        be.pad(cur)
    return fn


def cpu_itdq(payload):
    from oracle.pyoracle import Reference
    lev, lw, lh, qp, bd, iqt = payload
    be = Reference(2)
    f = be.lib.ref_itdq_blocks
    f.restype = None
    f.argtypes = [C.c_void_p] + [C.c_int] * 6
    buf = lev.copy()

    def fn():
        buf[...] = lev
        f(buf.ctypes.data, len(lev), lw, lh, qp, bd, int(iqt))
    return fn


def cpu_mc(payload):
    from oracle.pyoracle import Reference
    plane, origin, stride, chroma, mv, w, h, bd = payload
    be = Reference(2)
    f = be.lib.ref_mc_blocks
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 5
    out = np.zeros(len(mv) * w * h + 64, np.int16)

    def fn():
        f(plane.ctypes.data + 2 * origin, stride, chroma, mv.ctypes.data, out.ctypes.data, len(mv), w, h, bd, 0)
    return fn


# ---------------------------------------------------------------------------------------------------------------------------------------
def run(torch, ctx, stream, dev, peak, log=lambda s: None, npic=6, reps=4, cores=None):
    from oracle.pyoracle import have_reference
    from xevd_b200 import synth
    from xevd_b200.device import Context
    cores = cores or len(os.sched_getaffinity(0))
    have_ref = have_reference()
    out = []
    bd = 10

    def upload_work(cl):
        return dict(cl=cl, cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev), first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev),
                    ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev), coef=torch.from_numpy(cl.coef.copy()).to(dev),
                    max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max()))

    def recon(c, prm, cur, refs, refs1, wk, has_intra):
        cl = wk["cl"]
        c.recon_frame_dev(prm, cur, refs, refs1, wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(), cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext),
                          wk["coef"].data_ptr(), cl.coef.size, has_intra=has_intra, max_cu_per_ctu=wk["max_cu"])

    def timed(fn, n=npic, r=reps):
        """median device time (us) of fn(i) over r rotations through n picture slots"""
        for i in range(n):
            fn(i)
        ctx.sync()
        ev = []
        for _ in range(r):
            for i in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(i); e1.record(stream)
                ev.append((e0, e1))
        ctx.sync()
        return 1e3 * float(np.median([a.elapsed_time(b) for a, b in ev]))

    def entry(name, us, alg, what, cpu1=None, cpun=None, **more):
        gbs = alg / (us * 1e-6) / 1e9
        e = {"workload": name, "passes": what, "us_per_picture": round(us, 1), "frames_per_sec": round(1e6 / us, 1),
             "roofline": {"bound": "hbm", "algorithmic_bytes": int(alg), "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4)}}
        if cpu1 is not None:
            e["cpu_baseline"] = {"kind": "reference", "unit": "frames/s", "value_1_thread": round(cpu1, 2)}
            if cpun is not None:
                e["cpu_baseline"].update({"value_all_cores": round(cpun, 2), "cores": cores})
        e.update(more)
        out.append(e)
        log(f"extra: {name}: {us:.1f} us, {e['roofline']['frac']:.3f} of HBM peak" + (f", CPU 1 thread {cpu1:.1f}/s" if cpu1 else "") + (f", {cores} cores {cpun:.1f}/s" if cpun else ""))

    # ---- inter pictures --------------------------------------------------------------------------------------------------------
    def inter_case(name, w, h, variant, all_cores=False, concurrent=3, **kw):
        n_refs = 1 if variant == "A" else 2
        host_refs = synth.make_refs(w, h, bd, n_refs, seed=7)
        drefs = [ctx.pic_alloc(w, h).upload(r) for r in host_refs]
        curs = [ctx.pic_alloc(w, h) for _ in range(npic)]
        works = []
        for i in range(2):
            prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=1 + i, n_refs=n_refs, **kw)
            works.append((prm, cl, upload_work(cl)))
        r1 = [] if variant == "A" else drefs[::-1]
        us = timed(lambda i: (recon(ctx, works[i % 2][0], curs[i], drefs, r1, works[i % 2][2], 0), ctx.pad(curs[i])))
        more = {}
        if concurrent:
            # independent GOPs side by side, one context / stream each (what bench.py's headline leg does): the tail wave of one picture's grid
            # overlaps the head of another's - a 1080p picture is 510 CTUs, 1.15 waves of the 444 resident CTAs.  Device time between one event
            # all streams wait for and the last stream's end; enough picture slots to exceed the L2.
            pic_mb = (w + 288) * (h + 288) * 3 / 1e6
            per = max(2, int(np.ceil(160.0 / pic_mb / concurrent)))
            cs, sts, pics = [], [], []
            for k in range(concurrent):
                st = torch.cuda.Stream(device=dev)
                c = Context(dev.index)
                c.set_stream(st.cuda_stream)
                cs.append(c); sts.append(st); pics.append([c.pic_alloc(w, h) for _ in range(per)])

            def sweep(n):
                for j in range(n):
                    for k, c in enumerate(cs):
                        wk_ = works[(j + k) % 2]
                        recon(c, wk_[0], pics[k][j % per], drefs, r1, wk_[2], 0)
                        c.pad(pics[k][j % per])
            sweep(per)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sts[0])
            for st in sts[1:]:
                st.wait_event(e0)
            n_sweep = 4 * per
            sweep(n_sweep)
            for st in sts[1:]:
                d_ = torch.cuda.Event()
                d_.record(st)
                sts[0].wait_event(d_)
            e1.record(sts[0])
            torch.cuda.synchronize()
            more = {"gop_streams": concurrent, "frames_per_sec_concurrent": round(n_sweep * concurrent / (e0.elapsed_time(e1) * 1e-3), 1)}
            for k, c in enumerate(cs):
                for p_ in pics[k]:
                    p_.free()
                c.close()
        cpu1 = cpun = None
        if have_ref:
            payload = (works[0][0], works[0][1], host_refs, [] if variant == "A" else host_refs[::-1], None)
            cpu1 = cpu_rate("cpu_picture", payload, 1, 2)
            if all_cores:
                cpun = cpu_rate("cpu_picture", payload, cores, 2)
        entry(name, us, alg_recon(works[0][1]), "xb200_recon_frame_dev + xb200_pad", cpu1, cpun, picture=f"{w}x{h} 4:2:0 {bd}-bit", **more)
        for p in drefs + curs:
            p.free()

    inter_case("4k-2B", 3840, 2160, "B", all_cores=True)
    inter_case("4k-iqt", 3840, 2160, "B", iqt=True, main_mv=True)
    inter_case("1080p-2A", 1920, 1080, "A")
    inter_case("8k-2A", 7680, 4320, "A")

    # ---- pictures with the wavefront kernel, and config 3 ----------------------------------------------------------------------
    w, h = 3840, 2160
    host_refs = synth.make_refs(w, h, bd, 2, seed=7)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in host_refs]
    curs = [ctx.pic_alloc(w, h) for _ in range(npic)]

    def wavefront_case(name, prm, cl, refs0, refs1, hrefs0, hrefs1, concurrent=0, has=1):
        wk = upload_work(cl)
        us = timed(lambda i: (recon(ctx, prm, curs[i], refs0, refs1, wk, has), ctx.pad(curs[i])))
        more = {}
        if concurrent:       # independent pictures in flight on separate streams (contexts): what a GOP-parallel caller gets from one GPU
            cs, sts, pics = [], [], []
            for k in range(concurrent):
                st = torch.cuda.Stream(device=dev)
                c = Context(dev.index)
                c.set_stream(st.cuda_stream)
                cs.append(c); sts.append(st); pics.append([c.pic_alloc(w, h) for _ in range(2)])
            def sweep(n):
                for j in range(n):
                    for k, c in enumerate(cs):
                        recon(c, prm, pics[k][j & 1], refs0, refs1, wk, has)
                        c.pad(pics[k][j & 1])
                for c in cs:
                    c.sync()
            sweep(1)
            t0 = time.perf_counter()
            sweep(3)
            dt = time.perf_counter() - t0
            more = {"pictures_in_flight": concurrent, "frames_per_sec_concurrent": round(3 * concurrent / dt, 1)}
            for k, c in enumerate(cs):
                for p in pics[k]:
                    p.free()
                c.close()
        cpu1 = cpun = None
        if have_ref:
            payload = (prm, cl, hrefs0, hrefs1, None)
            cpu1 = cpu_rate("cpu_picture", payload, 1, 1)
            cpun = cpu_rate("cpu_picture", payload, cores, 1)
        entry(name, us, alg_recon(cl), "xb200_recon_frame_dev (inter kernels + CTU wavefront kernel) + xb200_pad", cpu1, cpun, picture=f"{w}x{h} 4:2:0 {bd}-bit", **more)

    prm_x, cl_x = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=21, n_refs=2, coded_frac=0.7)
    synth.add_intra_cus(cl_x, np.random.default_rng(6), 0.1)
    synth.derive_avail_cu(cl_x)
    wavefront_case("4k-P-intra10", prm_x, cl_x, drefs, drefs[::-1], host_refs, host_refs[::-1], concurrent=4)
    for eipd in (0, 1):
        prm_i, cl_i = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=9, n_refs=1, coded_frac=0.7, iqt=bool(eipd))
        prm_i.tool_eipd = prm_i.tool_htdf = eipd
        prm_i.slice_qp = 34
        synth.add_intra_cus(cl_i, np.random.default_rng(2), 1.0, eipd=bool(eipd))
        synth.derive_avail_cu(cl_i)
        wavefront_case("4k-I-eipd-htdf" if eipd else "4k-I-baseline", prm_i, cl_i, drefs[:1], [], host_refs[:1], [], concurrent=6, has=5)

    # config 3: everything on one Main picture
    prm_m, cl_m, refs_m, alf, flags = synth.make_main_frame(w, h, bit_depth=bd, seed=3)
    dm = [ctx.pic_alloc(w, h).upload(r) for r in refs_m]
    wk_m = upload_work(cl_m)
    ids = ((0, 1), (1, 0))
    ctx.set_chroma_qp_table(synth.chroma_qp_table(True))
    has = 3 if (cl_m.cus["flags"] & 3 != 3).any() else 1
    # what xb200_recon_frame derives from the work list itself (libxevd_b200.cu): more than half of the CUs on the wavefront -> persistent grid
    n_wave = int(((cl_m.cus["mode"] == 0) | (cl_m.cus["mode"] == 4) | ((cl_m.cus["cbf"] & 15) != 0)).sum())
    if 2 * n_wave > cl_m.n_cu:
        has |= 4
    parts = {}
    parts["recon"] = timed(lambda i: recon(ctx, prm_m, curs[i], dm, dm[::-1], wk_m, has))
    parts["deblock"] = timed(lambda i: ctx.deblock(prm_m, curs[i], dm, dm[::-1]))
    parts["alf"] = timed(lambda i: ctx.alf(prm_m, curs[i], alf, flags))
    parts["pad"] = timed(lambda i: ctx.pad(curs[i]))
    us = timed(lambda i: (recon(ctx, prm_m, curs[i], dm, dm[::-1], wk_m, has), ctx.deblock(prm_m, curs[i], dm, dm[::-1]), ctx.alf(prm_m, curs[i], alf, flags), ctx.pad(curs[i])))
    samples = w * h * 3 // 2
    nscu = (w // 4) * (h // 4)
    alg = alg_recon(cl_m) + (samples * 4 + nscu * 15) + samples * 4
    cpu1 = cpun = None
    if have_ref:
        payload = (prm_m, cl_m, refs_m, refs_m[::-1], (alf, flags, ids))
        cpu1 = cpu_rate("cpu_picture", payload, 1, 1)
        cpun = cpu_rate("cpu_picture", payload, cores, 1)
    # the same pipeline with four independent pictures in flight (separate contexts / streams): the wavefront kernel keeps a fraction of the SMs
    # busy, so pictures that do not reference each other overlap
    conc = 4
    cs, pics = [], []
    for k in range(conc):
        st = torch.cuda.Stream(device=dev)
        c = Context(dev.index)
        c.set_stream(st.cuda_stream)
        c.set_chroma_qp_table(synth.chroma_qp_table(True))
        cs.append((c, st)); pics.append([c.pic_alloc(w, h) for _ in range(2)])
    def sweep_m(n):
        for j in range(n):
            for k, (c, _) in enumerate(cs):
                cur = pics[k][j & 1]
                recon(c, prm_m, cur, dm, dm[::-1], wk_m, has); c.deblock(prm_m, cur, dm, dm[::-1]); c.alf(prm_m, cur, alf, flags); c.pad(cur)
        for c, _ in cs:
            c.sync()
    sweep_m(1)
    t0 = time.perf_counter()
    sweep_m(3)
    fps_conc = 3 * conc / (time.perf_counter() - t0)
    for k, (c, _) in enumerate(cs):
        for p in pics[k]:
            p.free()
        c.close()
    entry("4k-main-full", us, alg, "xb200_recon_frame_dev (all Main tools) + xb200_deblock (ADDB) + xb200_alf + xb200_pad: BASELINE config 3", cpu1, cpun,
          picture=f"{w}x{h} 4:2:0 {bd}-bit", us_by_call={k: round(v, 1) for k, v in parts.items()}, pictures_in_flight=conc, frames_per_sec_concurrent=round(fps_conc, 1))
    ctx.set_chroma_qp_table(synth.chroma_qp_table(False))
    for p in dm + drefs + curs:
        p.free()

    # ---- config 5: leaf kernels ------------------------------------------------------------------------------------------------
    n_s = 1 << 22
    rows = []
    for iqt in (0, 1):
        for lw in range(1, 7):
            for lh in range(1, 7):
                nb = n_s >> (lw + lh)
                rng = np.random.default_rng(10 * lw + lh)
                base = synth.quantised_dct(rng.laplace(0, 8.0, (64, 1 << lh, 1 << lw)), 44, bool(iqt))
                if iqt and lw == 6:
                    base[:, :, 32:] = 0
                if iqt and lh == 6:
                    base[:, 32:, :] = 0
                lev = np.ascontiguousarray(np.tile(base, ((nb + 63) // 64, 1, 1))[:nb])
                d_in = torch.from_numpy(lev.reshape(-1)).to(dev)
                d_out = torch.empty_like(d_in)
                us = timed(lambda i: ctx.itdq_blocks_dev(d_in.data_ptr(), d_out.data_ptr(), nb, lw, lh, 44, bd, bool(iqt)), n=1, r=5)
                row = {"shape": f"{1 << lw}x{1 << lh}", "iqt": iqt, "us": round(us, 1), "gbs": round(4 * n_s / us / 1e3, 1), "frac": round(4 * n_s / us / 1e3 / peak, 4)}
                if have_ref and lw == lh:
                    ncpu = min(nb, 1 << 12)
                    row["cpu_msamples_per_s_1_thread"] = round(cpu_rate("cpu_itdq", (lev[:ncpu].copy(), lw, lh, 44, bd, iqt), 1, 3) * (ncpu << (lw + lh)) / 1e6, 1)
                    row["gpu_msamples_per_s"] = round(n_s / us, 1)
                rows.append(row)
    out.append({"workload": "config5-itdq", "passes": "xb200_itdq_blocks_dev: dequant + 2-D inverse transform of 4 Mi samples per launch, every (w, h) in 2..64, Baseline and IQT; "
                "algorithmic bytes 4 B per sample (s16 in, s16 out)", "rows": rows})
    log("extra: config5-itdq: " + ", ".join(f"{r['shape']}{'q' if r['iqt'] else ''} {r['us']}us" for r in rows if r["shape"] in ("4x4", "16x16", "64x64", "64x8")))
    W, H = 3840, 2160
    from xevd_b200.frame import HostPicture
    pic = HostPicture.random(W, H, bd, np.random.default_rng(5))
    pic.pad_borders()
    dpic = ctx.pic_alloc(W, H).upload(pic)
    rows = []
    for chroma in (0, 1):
        for s in (8, 16, 32, 64):
            bw = s >> chroma
            pw, ph = W >> chroma, H >> chroma
            sh = 5 if chroma else 4
            nb = (1 << 22) // (bw * bw)
            rng = np.random.default_rng(s + chroma)
            x = rng.integers(-16, pw - bw + 16, nb); y = rng.integers(-16, ph - bw + 16, nb)
            fx = rng.integers(1, 4, nb) * (1 << (sh - 2)); fy = rng.integers(1, 4, nb) * (1 << (sh - 2))      # quarter-pel phases, both fractional: variant nn
            mv = np.stack([(x << sh) + fx, (y << sh) + fy, fx, fy], 1).astype(np.int32)
            d_mv = torch.from_numpy(mv).to(dev)
            d_out = torch.zeros(nb * bw * bw, dtype=torch.int16, device=dev)
            us = timed(lambda i: ctx.mc_blocks_dev(dpic, 1 if chroma else 0, d_mv.data_ptr(), d_out.data_ptr(), nb, bw, bw, bd, False), n=1, r=5)
            alg = 4 * nb * bw * bw + 16 * nb         # one reference sample read + one prediction sample written per output sample, + the vectors
            row = {"plane": "chroma 4-tap" if chroma else "luma 8-tap", "block": f"{bw}x{bw}", "us": round(us, 1), "gbs": round(alg / us / 1e3, 1), "frac": round(alg / us / 1e3 / peak, 4),
                   "gpu_msamples_per_s": round(nb * bw * bw / us, 1)}
            if have_ref:
                ncpu = min(nb, 1 << 12)
                plane = pic.buf_u if chroma else pic.buf_y
                stride = plane.shape[1]
                pad = pic.pad_c if chroma else pic.pad_l
                row["cpu_msamples_per_s_1_thread"] = round(cpu_rate("cpu_mc", (plane, pad * stride + pad, stride, chroma, mv[:ncpu].copy(), bw, bw, bd), 1, 3) * ncpu * bw * bw / 1e6, 1)
            rows.append(row)
    dpic.free()
    out.append({"workload": "config5-mc", "passes": "xb200_mc_blocks_dev: variant nn (both phases fractional) on 4 Mi samples per launch from a 4K reference plane", "rows": rows})
    log("extra: config5-mc: " + ", ".join(f"{r['plane'][:1]}{r['block']} {r['us']}us" for r in rows))

    # ---- end to end with 8-bit output: the same public calls as bench.py's e2e leg on the headline workload, the decoded picture pulled as 8-bit
    #      planes (what xevd_app --output-bit-depth 8 asks for; xb200_pic_pull narrows on the device, xevd_app_util.h:359-381): half the D2H
    #      bytes of the headline's e2e, which sits on that copy
    try:
        import ctypes as C
        from xevd_b200.frame import sparse_coef
        w8, h8, bd8, F8, nctx = 3840, 2160, 10, 12, 3
        work = []
        for i in range(4):
            prm8, cl8 = synth.make_inter_frame(w8, h8, bit_depth=bd8, variant="A", seed=100 + i, n_refs=1)
            ent, cf = sparse_coef(cl8.coef)
            work.append(dict(prm=prm8, cl=cl8, cus=torch.from_numpy(cl8.cus.view(np.uint8).copy()).pin_memory(),
                             first=torch.from_numpy(cl8.ctu_first.view(np.int32).copy()).pin_memory(), ext=torch.from_numpy(cl8.ext.view(np.uint8).copy()).pin_memory(),
                             ent=torch.from_numpy(ent.view(np.int32).copy()).pin_memory(), cf=torch.from_numpy(cf.view(np.int32).copy()).pin_memory()))
        refs8 = synth.make_refs(w8, h8, bd8, 1, seed=8)
        cs = []
        for k in range(nctx):
            st = torch.cuda.Stream(device=dev)
            c = Context(dev.index)
            c.set_stream(st.cuda_stream)
            cs.append(dict(c=c, st=st, ref=c.pic_alloc(w8, h8).upload(refs8[0])))
        slots8 = []
        for i in range(F8):
            c = cs[i % nctx]
            slots8.append(dict(c=c, cur=c["c"].pic_alloc(w8, h8), wk=work[i % 4], y=torch.empty((h8, w8), dtype=torch.uint8).pin_memory(),
                               u=torch.empty((h8 // 2, w8 // 2), dtype=torch.uint8).pin_memory(), v=torch.empty((h8 // 2, w8 // 2), dtype=torch.uint8).pin_memory()))
        def step8():
            for s8 in slots8:
                c, wk = s8["c"]["c"], s8["wk"]
                cl8 = wk["cl"]
                rh = (C.c_void_p * 1)(s8["c"]["ref"].handle)
                c._chk(c.lib.xb200_recon_frame_sparse(c.handle, C.byref(wk["prm"]), s8["cur"].handle, rh, 1, rh, 0, wk["cus"].data_ptr(), cl8.n_cu,
                                                      wk["first"].data_ptr(), cl8.n_ctu, wk["ext"].data_ptr(), len(cl8.ext), wk["ent"].data_ptr(), wk["ent"].numel(),
                                                      wk["cf"].data_ptr(), cl8.coef.size), "xb200_recon_frame_sparse")
                c.pad(s8["cur"])
                c._chk(c.lib.xb200_pic_pull(c.handle, s8["cur"].handle, None, 8, 0, 0, 0, 0, s8["y"].data_ptr(), w8, s8["u"].data_ptr(), w8 // 2,
                                            s8["v"].data_ptr(), w8 // 2), "xb200_pic_pull")
            for k in cs:
                k["c"].sync()
        step8()
        t0 = time.perf_counter()
        for _ in range(4):
            step8()
        fps8 = 4 * F8 / (time.perf_counter() - t0)
        h2d8 = int(np.mean([wk["cus"].numel() + wk["first"].numel() * 4 + wk["ext"].numel() + wk["ent"].numel() * 4 + wk["cf"].numel() * 4 for wk in work]))
        out.append({"workload": "4k-2A-e2e-8bit-output", "passes": "xb200_recon_frame_sparse (host buffers) + xb200_pad + xb200_pic_pull to 8-bit planes, 3 contexts / streams: the e2e leg of "
                    "bench.py with the decoded 10-bit picture delivered as 8-bit output", "frames_per_sec": round(fps8, 1), "h2d_bytes_per_picture": h2d8, "d2h_bytes_per_picture": w8 * h8 * 3 // 2,
                    "checksum": int(slots8[0]["y"][::64, ::64].to(torch.int64).sum().item()), "picture": f"{w8}x{h8} 4:2:0 10-bit, 8-bit output"})
        log(f"extra: 4k-2A-e2e-8bit-output: {fps8:.0f} pictures/s")
        for s8 in slots8:
            s8["cur"].free()
        for k in cs:
            k["ref"].free(); k["c"].close()
    except Exception as e:
        out.append({"workload": "4k-2A-e2e-8bit-output", "error": repr(e)})

    # ---- whole-decoder drop-in on real elementary streams --------------------------------------------------------------------
    try:
        from xevd_b200 import xevd_api as X
        if X.GPU_SO.exists():
            for name in ("base_1080p_8b", "main_1080p_10b"):
                path = ROOT / "tests" / "golden" / "streams" / f"{name}.evc"
                if not path.exists():
                    continue
                nals = X.read_stream(path)
                res = {}
                for tag, so in (("gpu", X.GPU_SO), ("reference_cpu_1_thread", X.REF_SO)):
                    if not so.exists():
                        continue
                    lib = X.XevdLibrary(so)
                    X.decode_stream(lib, nals)                  # warm-up (first use of the library: CUDA context, code load)
                    whole, per_pic = [], []
                    for _ in range(3):
                        with X.Decoder(lib) as d:               # xevd_create / xevd_delete outside the timed region
                            # the stream is fed three times to the same instance (it starts with an IDR picture): the first pass pays the
                            # sequence set-up (picture buffers, page-locking, staging), passes two and three are the steady state
                            for rep in range(3):
                                t_all, n = 0.0, 0
                                for nal in nals:
                                    t0 = time.perf_counter()
                                    ret, stat = d.decode(nal)
                                    got = 0
                                    while d.pull() is not None:
                                        got += 1
                                    dt = time.perf_counter() - t0
                                    t_all += dt
                                    n += got
                                    if got and rep > 0:
                                        per_pic.append(dt / got)
                                if rep == 0:
                                    whole.append(n / t_all)
                    res[tag] = {"first_pass_incl_setup": round(float(np.median(whole)), 1), "steady_state": round(1.0 / float(np.median(per_pic)), 1)}
                out.append({"workload": f"stream-{name}", "passes": "xevd_create / xevd_decode / xevd_pull on a generated elementary stream: entropy decoding and motion derivation on ONE host "
                            "thread in both libraries (the reference's own code); libxevd_gpu.so reconstructs on the device and copies every picture back. first_pass_incl_setup is the first pass over the 6..8-picture stream "
                            "with the one-time set-up of a decoder instance (device pictures, page-locked planes, staging ring, ALF / wavefront scratch), steady_state is 1 / median decode+pull time per picture when the same instance "
                            "decodes the stream a second and a third time",
                            "frames_per_sec": res})
                log(f"extra: stream-{name}: {res}")
    except Exception as e:        # the drop-in library is optional at bench time
        out.append({"workload": "stream", "error": repr(e)})
    return out
