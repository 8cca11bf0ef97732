#!/usr/bin/env python3
"""bench.py -- headline benchmark of the XEVD reconstruction hot path on B200.

Metric (BASELINE.json): 4K 10-bit frames/s of the MC + ITDQ + recon path on synthetic pre-parsed CU arrays,
with the achieved fraction of the HBM roofline, next to the reference's own CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 4k-2A|...]

A "step" is one batch of FRAMES_PER_STEP distinct pictures (reference picture, CU array, coefficient stream and
output picture all distinct per slot, ~85 MB each at 4K, so a step streams > 10x the 126 MB L2).  Per picture the
step runs xb200_recon_frame_dev (the picture-level replacement of xevd_ctu_row_rec_mt) followed by xb200_pad
(xevd_picbuf_expand).  `value` times that with every input already resident in HBM; `e2e` times the same pictures
through the host-buffer C ABI call xb200_recon_frame with pinned host inputs (H2D inside the timed region) and a
D2H of every decoded picture.

Under torchrun each rank drives its own GPU on its own share of pictures (GOP-level sharding: no data-path
collective); the timed region is bracketed by barriers and the max over ranks is reported.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (w, h, bit_depth, variant)
    "4k-2A": (3840, 2160, 10, "A"),
    "4k-2B": (3840, 2160, 10, "B"),
    "1080p-2A": (1920, 1080, 10, "A"),
    "1080p-2B": (1920, 1080, 10, "B"),
    "8k-2A": (7680, 4320, 10, "A"),
}
REF_SO_PATH = ROOT / "oracle" / "_ref" / "libxevd_ref.so"
METRIC = "4k_10bit_frames_per_sec"
UNIT = "frames/s"


def algorithmic_bytes(w, h, cl, prm_bi_frac=0.0):
    """compulsory traffic, each datum once: reference read 2 B/sample per prediction direction,
    coefficients 2 B per coded sample, reconstruction write 2 B/sample, plus the CU descriptors."""
    cus = cl.cus
    samples = (1 << (cus["log2w"].astype(np.int64) + cus["log2h"].astype(np.int64))) * 3 // 2
    ndir = (cus["refi"][:, 0] >= 0).astype(np.int64) + (cus["refi"][:, 1] >= 0).astype(np.int64)
    ref = int((samples * ndir).sum()) * 2
    coef = int(cl.coef.size) * 2
    rec = int(samples.sum()) * 2
    desc = cus.nbytes + cl.ctu_first.nbytes
    return ref + coef + rec + desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def recorded_counters(workload):
    """what only a profiler can count, from the newest committed ncu capture of the dominant kernel on this workload (profiles/rN/traffic_*.json):
    DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), warp instructions per launch, issue-slot utilisation"""
    best = {}
    for p in sorted((ROOT / "profiles").glob("r*/traffic_*.json")):
        try:
            d = json.loads(p.read_text())
        except (OSError, ValueError):
            continue
        if d.get("workload") == workload:
            best = dict(d, file=str(p.relative_to(ROOT)))
    return best


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def common_config(args, world):
    """the `config` object: identical in both arms (--impl ours / reference), so that the driver's same-config check sees one workload"""
    w, h, bd, variant = WORKLOADS[args.workload]
    return {"workload": args.workload, "picture": f"{w}x{h} 4:2:0 {bd}-bit",
            "cu_partition": "uniform 16x16 uni-pred all-coded" if variant == "A" else "quadtree 64..8, 50% bi-pred",
            "per_picture": "MC + dequant / inverse transform + reconstruction of every CU (xevd_ctu_row_rec_mt), then border padding (xevd_picbuf_expand)",
            "sharding": args.sharding, "n_gpus": world,
            "per_device": "independent GOPs decoded side by side (GPU arm: three streams per GPU; reference arm: one decoder instance per host core)",
            "l2": "every step streams more than the 126 MB L2 (distinct pictures in rotation)"}


def make_workload(name, n_distinct, seed0=1):
    from xevd_b200 import synth
    w, h, bd, variant = WORKLOADS[name]
    frames = []
    for i in range(n_distinct):
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=seed0 + i, n_refs=1 if variant == "A" else 2)
        frames.append((prm, cl))
    return w, h, bd, variant, frames


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, dispatched AVX2 kernels:
    xevd_mc + xevdm_sub_block_itdq + xevd_recon per CU, then xevd_picbuf_expand), one single-threaded decoder
    instance per host core, each on its own pictures (GOP-level parallelism, the only way the reference scales
    past XEVD_MAX_TASK_CNT).  A step is a bounded sample: one picture per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle.pyoracle import have_reference
    kind = "reference" if have_reference() else "port"
    if kind == "reference":
        C.CDLL(str(REF_SO_PATH))          # also in the parent, so that the driver's loaded-library record shows the reference (the workers are forks)
    cores = len(os.sched_getaffinity(0))
    w, h, bd, variant, frames = make_workload(args.workload, 1)
    prm, cl = frames[0]
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    start = ctx.Barrier(cores)

    def worker(idx):
        from oracle.pyoracle import Oracle, Reference
        from xevd_b200 import synth
        from xevd_b200.frame import HostPicture
        be = Reference(2) if kind == "reference" else Oracle()
        refs = synth.make_refs(w, h, bd, 1 if variant == "A" else 2, seed=50 + idx)
        cur = HostPicture(w, h, prm.poc)
        for _ in range(args.warmup):
            be.recon_frame(prm, cur, refs, refs[::-1], cl); be.pad(cur)
        start.wait()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            be.recon_frame(prm, cur, refs, refs[::-1], cl); be.pad(cur)
        q.put(time.perf_counter() - t0)

    procs = [ctx.Process(target=worker, args=(i,)) for i in range(cores)]
    for p in procs:
        p.start()
    times = [q.get() for _ in procs]
    for p in procs:
        p.join()
    elapsed = max(times)
    fps = cores * args.steps / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s16", "data": "synthetic",
        "config": common_config(args, args.gpus),
        "detail": {"frames_per_step": cores, "native_library": str(REF_SO_PATH)},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{cores} single-threaded decoder instances x {args.steps} pictures of {args.workload} each (recon + pad)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def cpu_baseline_sample(workload, seconds_budget=20.0):
    """bounded CPU sample on rank 0: the reference (or the oracle port) single-threaded on whole pictures"""
    from oracle.pyoracle import Oracle, Reference, have_reference
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    kind = "reference" if have_reference() else "port"
    be = Reference(2) if kind == "reference" else Oracle()
    w, h, bd, variant, frames = make_workload(workload, 1)
    prm, cl = frames[0]
    refs = synth.make_refs(w, h, bd, 1 if variant == "A" else 2, seed=77)
    cur = HostPicture(w, h, prm.poc)
    be.recon_frame(prm, cur, refs, refs[::-1], cl)       # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        be.recon_frame(prm, cur, refs, refs[::-1], cl); be.pad(cur)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds_budget or n >= 50:
            break
    out = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{n} pictures of {workload} (recon + pad), 1 thread, {dt:.1f} s"}
    if kind == "reference":       # the 8-thread figure next to the single-thread one (8 single-threaded instances, GOP-parallel)
        from tools import bench_extra
        out["value_8_threads"] = bench_extra.cpu_rate("cpu_picture", (prm, cl, refs, refs[::-1], None), 8, 3)
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from xevd_b200.device import Context
    from xevd_b200.frame import HostPicture
    from xevd_b200 import synth

    from xevd_b200 import dist as xdist
    rank, world, local = xdist.init("nccl")        # GOP-level sharding: one process per GPU, no data-path collective
    dist = None
    if world > 1:
        import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    F = args.slots                                         # distinct picture slots (> L2 per rotation)
    P = max(1, args.frames_per_step // F)                  # rotations per step: frames per step = P * F
    band = args.sharding in ("band", "band-p2p")
    p2p = args.sharding == "band-p2p"
    # GOP sharding: every rank has its own pictures.  Band sharding: every rank works on the SAME pictures, one CTU-row band each.
    w, h, bd, variant, frames = make_workload(args.workload, min(4, F), seed0=1 + (0 if band else 10 * rank))
    if band:
        bands = xdist.band_partition(h, 6, world)
        r0, k = bands[rank]
        banded = []
        for prm, cl in frames:
            bp = type(prm).from_buffer_copy(prm)
            bp.ctu_row0, bp.ctu_rows = r0, k
            banded.append((bp, cl.band(r0, k)))
        full_frames, frames = frames, banded
    n_refs = 1 if variant == "A" else 2

    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local)
    ctx.set_stream(stream.cuda_stream)
    # GOP sharding: pictures of different GOPs do not depend on each other, so one GPU decodes G GOPs side by side, one context / stream per
    # GOP (the reference arm's host cores do the same: one decoder instance per core).  The kernels of one stream follow each other; the last
    # wave of a picture's CTAs (2040 CTUs over 444 resident CTAs = 4.6 waves) shares the SMs with the first wave of the other GOP's picture.
    n_gop = 1 if band else max(1, int(os.environ.get("XB200_BENCH_GOP_STREAMS", "3")))
    gop_ctx, gop_stream = [ctx], [stream]
    for _ in range(n_gop - 1):
        st_ = torch.cuda.Stream(device=dev)
        c_ = Context(local)
        c_.set_stream(st_.cuda_stream)
        gop_ctx.append(c_); gop_stream.append(st_)

    # ---- resident inputs: F slots, each with its own reference picture(s), CU array, coefficients, output picture
    host_refs = synth.make_refs(w, h, bd, 2, seed=1000 + (0 if band else rank))
    slots = []
    with torch.cuda.stream(stream):
        for i in range(F):
            prm, cl = frames[i % len(frames)]
            sc = gop_ctx[i % n_gop]                      # a device picture belongs to the context / stream that fills it
            refs = [sc.pic_alloc(w, h).upload(host_refs[(i + j) % 2]) for j in range(n_refs)]
            for j, r in enumerate(refs):
                r.set_poc(host_refs[(i + j) % 2].poc if n_refs > 1 else 0)
            cur = sc.pic_alloc(w, h)
            d_cus = torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev)
            d_first = torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev)
            d_ext = torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev)
            d_coef = torch.from_numpy(cl.coef.copy()).to(dev)
            slots.append(dict(ctx=sc, prm=prm, cl=cl, refs=refs, refs_l1=(refs[::-1] if variant != "A" else []), cur=cur, d_cus=d_cus, d_first=d_first, d_ext=d_ext, d_coef=d_coef,
                              max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max())))
    torch.cuda.synchronize()
    exch = xdist.BandExchange(ctx, slots[0]["cur"], 6, rank, world, dev) if (band and not p2p) else None
    flag = torch.zeros(1, device=dev)
    if p2p:
        for s in slots:
            xdist.open_peer_pictures(ctx, s["cur"])       # once per picture buffer: map the other ranks' copies (CUDA IPC)
        torch.cuda.synchronize()
        xdist.barrier()

    def stream_barrier():
        if dist is not None:
            with torch.cuda.stream(stream):
                dist.all_reduce(flag)                     # tiny collective in stream order: all peer stores of this picture have landed

    def step_resident():
        for _ in range(P):                                 # a step = P rotations over the F distinct picture slots
            for s in slots:
                cl = s["cl"]
                if cl.n_cu:
                    s["ctx"].recon_frame_dev(s["prm"], s["cur"], s["refs"], s["refs_l1"], s["d_cus"].data_ptr(), cl.n_cu,
                                             s["d_first"].data_ptr(), cl.n_ctu, s["d_ext"].data_ptr(), len(cl.ext), s["d_coef"].data_ptr(), cl.coef.size,
                                             max_cu_per_ctu=s["max_cu"])
                if exch is not None:
                    with torch.cuda.stream(stream):
                        exch.exchange(s["cur"])              # one in-place NCCL all-gather of the packed bands per picture
                elif p2p:
                    stream_barrier()
                s["ctx"].pad(s["cur"])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ----------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sum(c_.launches for c_ in gop_ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for st_ in gop_stream[1:]:
            st_.wait_event(e0)                            # the timed region opens on every GOP stream at e0 ...
        for _ in range(args.steps):
            step_resident()
        for st_ in gop_stream[1:]:
            done_ = torch.cuda.Event()
            done_.record(st_)
            stream.wait_event(done_)                      # ... and closes when the last of them has finished
        e1.record(stream)
    barrier()
    launches = sum(c_.launches for c_ in gop_ctx) - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    ms = xdist.max_over_ranks(ms, dev)
    fps = (1 if band else world) * P * F * args.steps / (ms * 1e-3)

    # ---- the same loop on ONE stream (the first GOP stream's own pictures, still more than the L2 in rotation): the figure the three-stream
    #      `value` is to be read against; launches x kernel_ms fits into this one, not into the overlapped step
    single_fps = None
    if n_gop > 1:
        own_ = [s for s in slots if s["ctx"] is ctx]
        n_rot = max(1, (P * F * min(args.steps, 20)) // len(own_))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            s0.record(stream)
            for _ in range(n_rot):
                for s in own_:
                    cl = s["cl"]
                    ctx.recon_frame_dev(s["prm"], s["cur"], s["refs"], s["refs_l1"], s["d_cus"].data_ptr(), cl.n_cu,
                                        s["d_first"].data_ptr(), cl.n_ctu, s["d_ext"].data_ptr(), len(cl.ext), s["d_coef"].data_ptr(), cl.coef.size,
                                        max_cu_per_ctu=s["max_cu"])
                    ctx.pad(s["cur"])
            s1.record(stream)
        barrier()
        single_fps = world * n_rot * len(own_) / (xdist.max_over_ranks(s0.elapsed_time(s1), dev) * 1e-3)

    # ---- roofline of the dominant kernel (k_recon_inter): per-launch CUDA-event timing on the launching stream --------
    own = [s for s in slots if s["ctx"] is ctx]           # the pictures of the first GOP stream, one kernel at a time (> L2 in rotation)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(len(own) * min(args.steps, 4) * n_gop)]
    k = 0
    with torch.cuda.stream(stream):
        for _ in range(min(args.steps, 4) * n_gop):
            for s in own:
                cl = s["cl"]
                ev[k][0].record(stream)
                ctx.recon_frame_dev(s["prm"], s["cur"], s["refs"], s["refs_l1"], s["d_cus"].data_ptr(), cl.n_cu,
                                    s["d_first"].data_ptr(), cl.n_ctu, s["d_ext"].data_ptr(), len(cl.ext), s["d_coef"].data_ptr(), cl.coef.size,
                                max_cu_per_ctu=s["max_cu"])
                ev[k][1].record(stream)
                ctx.pad(s["cur"])
                k += 1
    torch.cuda.synchronize()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    alg = float(np.mean([algorithmic_bytes(w, h, s["cl"]) for s in slots]))       # band mode: this rank's band
    peak, peak_src = measured_peak()
    achieved = alg / (kern_ms * 1e-3) / 1e9
    counters = recorded_counters(args.workload)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------------------------------------
    n_ctx = 1 if band else int(os.environ.get("XB200_BENCH_CONTEXTS", "3"))          # band mode: the exchange orders every picture on one stream
    ctxs, streams = [], []
    for i in range(n_ctx):
        st = stream if band else torch.cuda.Stream(device=dev)
        c = ctx if band else Context(local)
        c.set_stream(st.cuda_stream)
        ctxs.append(c); streams.append(st)
    from xevd_b200.frame import sparse_coef
    pinned = []
    for i in range(F):
        prm, cl = frames[i % len(frames)]
        sp_entries, sp_first = sparse_coef(cl.coef)
        pc = dict(entries=torch.from_numpy(sp_entries.view(np.int32).copy()).pin_memory(), chunk_first=torch.from_numpy(sp_first.view(np.int32).copy()).pin_memory(),
                  cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).pin_memory(),
                  first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).pin_memory(),
                  ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).pin_memory(),
                  coef=torch.from_numpy(cl.coef.copy()).pin_memory(),
                  out_y=torch.empty((h, w), dtype=torch.int16).pin_memory(),
                  out_u=torch.empty((h // 2, w // 2), dtype=torch.int16).pin_memory(),
                  out_v=torch.empty((h // 2, w // 2), dtype=torch.int16).pin_memory())
        pinned.append(pc)
    # pictures owned per context (a device picture belongs to the stream that fills it)
    e2e_slots = []
    for i in range(F):
        if band:        # same pictures as the resident leg (their peer mappings / exchange buffers already exist)
            e2e_slots.append(dict(ctx=ctx, refs=slots[i]["refs"], cur=slots[i]["cur"]))
            continue
        c = ctxs[i % n_ctx]
        refs = [c.pic_alloc(w, h).upload(host_refs[(i + j) % 2]) for j in range(n_refs)]
        e2e_slots.append(dict(ctx=c, refs=refs, cur=c.pic_alloc(w, h)))
    torch.cuda.synchronize()
    h2d_dense = P * int(np.sum([p["cus"].numel() + p["first"].numel() * 4 + p["ext"].numel() + p["coef"].numel() * 2 for p in pinned]))
    h2d_sparse = P * int(np.sum([p["cus"].numel() + p["first"].numel() * 4 + p["ext"].numel() + p["entries"].numel() * 4 + p["chunk_first"].numel() * 4 for p in pinned]))
    d2h = P * F * (w * h * 3 // 2) * 2 if (not band or rank == 0) else 0       # band mode: rank 0 hands the assembled pictures to the consumer

    def step_e2e(sparse=True):
        for _ in range(P):
            for i in range(F):
                prm, cl = frames[i % len(frames)]
                s, p = e2e_slots[i], pinned[i]
                c = s["ctx"]
                if cl.n_cu and sparse:
                    # the coefficient stream crosses PCIe as (position, level) entries and is expanded on the device
                    c._chk(c.lib.xb200_recon_frame_sparse(c.handle, C.byref(prm), s["cur"].handle,
                                                          (C.c_void_p * n_refs)(*[r.handle for r in s["refs"]]), n_refs,
                                                          (C.c_void_p * n_refs)(*[r.handle for r in s["refs"][::-1]]), (n_refs if variant != "A" else 0),
                                                          p["cus"].data_ptr(), cl.n_cu, p["first"].data_ptr(), cl.n_ctu,
                                                          p["ext"].data_ptr(), len(cl.ext), p["entries"].data_ptr(), p["entries"].numel(),
                                                          p["chunk_first"].data_ptr(), cl.coef.size), "xb200_recon_frame_sparse")
                elif cl.n_cu:
                    c._chk(c.lib.xb200_recon_frame(c.handle, C.byref(prm), s["cur"].handle,
                                                   (C.c_void_p * n_refs)(*[r.handle for r in s["refs"]]), n_refs,
                                                   (C.c_void_p * n_refs)(*[r.handle for r in s["refs"][::-1]]), (n_refs if variant != "A" else 0),
                                                   p["cus"].data_ptr(), cl.n_cu, p["first"].data_ptr(), cl.n_ctu,
                                                   p["ext"].data_ptr(), len(cl.ext), p["coef"].data_ptr(), cl.coef.size), "xb200_recon_frame")
                if exch is not None:
                    with torch.cuda.stream(stream):
                        exch.exchange(s["cur"])
                elif p2p:
                    stream_barrier()
                c.pad(s["cur"])
                if band and rank != 0:
                    continue
                c._chk(c.lib.xb200_pic_download(c.handle, s["cur"].handle, p["out_y"].data_ptr(), w, p["out_u"].data_ptr(), w // 2,
                                                p["out_v"].data_ptr(), w // 2), "xb200_pic_download")
        for c in ctxs:
            c.sync()

    e2e_steps = max(1, min(args.steps, 6))

    def time_e2e(sparse):
        step_e2e(sparse)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e(sparse)
        torch.cuda.synchronize()
        return (1 if band else world) * P * F * e2e_steps / xdist.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_dense_fps = time_e2e(False)          # the dense coefficient stream over PCIe (xb200_recon_frame): what round 1 reported
    e2e_fps = time_e2e(True)                 # the sparse form (xb200_recon_frame_sparse): the public call a caller behind PCIe makes
    h2d = h2d_sparse
    # a decoded sample read back on the host proves the D2H happened
    checksum = int(pinned[0]["out_y"][::64, ::64].to(torch.int64).sum().item())

    # ---- north_star's multi-GPU data plane: rank 0 SCATTERS the pre-parsed CU work of one picture per rank over NCCL / NVLink, every rank
    #      reconstructs its picture, rank 0 GATHERS the decoded pictures (packed planes + per-SCU maps).  Everything stays on the devices;
    #      rank 0's links carry (world - 1) work lists out and (world - 1) pictures in per round, so this is the hub-limited figure next
    #      to the independent-GOP figure above (where every rank feeds itself).
    scatter_gather = None
    if dist is not None and not band:
        prm0, cl0 = frames[0]
        blob_parts = [cl0.cus.view(np.uint8).ravel(), cl0.ctu_first.view(np.uint8).ravel(), cl0.ext.view(np.uint8).ravel(), cl0.coef.view(np.uint8).ravel()]
        offs, o = [], 0
        for p_ in blob_parts:
            offs.append(o)
            o += (p_.size + 255) & ~255
        blob = np.zeros(o, np.uint8)
        for p_, of_ in zip(blob_parts, offs):
            blob[of_:of_ + p_.size] = p_
        recv = torch.empty(o, dtype=torch.uint8, device=dev)
        # rank 0 holds every rank's work list (here: the same synthetic picture shape, its own copy per destination)
        src = [torch.from_numpy(blob).to(dev) for _ in range(world)] if rank == 0 else None
        cur_sg = slots[0]["cur"]
        pic_bytes = ctx.band_bytes(cur_sg, h)
        mine = torch.empty(pic_bytes, dtype=torch.uint8, device=dev)
        gathered = [torch.empty(pic_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None

        def round_sg():
            with torch.cuda.stream(stream):
                dist.scatter(recv, src, src=0)
                b = recv.data_ptr()
                ctx.recon_frame_dev(prm0, cur_sg, slots[0]["refs"], slots[0]["refs_l1"], b + offs[0], cl0.n_cu, b + offs[1], cl0.n_ctu, b + offs[2], len(cl0.ext),
                                    b + offs[3], cl0.coef.size, max_cu_per_ctu=slots[0]["max_cu"])
                ctx.pad(cur_sg)
                ctx.band_pack(cur_sg, 0, h, mine.data_ptr())
                dist.gather(mine, gathered, dst=0)
        for _ in range(3):
            round_sg()
        barrier()
        n_rounds = 40
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(n_rounds):
            round_sg()
        g1.record(stream)
        barrier()
        ms_sg = xdist.max_over_ranks(g0.elapsed_time(g1), dev)
        scatter_gather = {"value": world * n_rounds / (ms_sg * 1e-3), "unit": UNIT, "rounds": n_rounds, "bytes_scattered_per_picture": int(o), "bytes_gathered_per_picture": int(pic_bytes),
                          "what": "per round: NCCL scatter of one picture's CU work per rank from rank 0, xb200_recon_frame_dev + xb200_pad + pack on every rank, NCCL gather of "
                                  "the decoded pictures (planes + maps) on rank 0; device-resident on both ends"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(args.workload)

    extra = None
    if rank == 0 and world == 1 and not band and not args.no_extra:
        # every other workload DESIGN.md quotes, measured in this same run (tools/bench_extra.py): per-picture device time, roofline fraction,
        # and the same pass through the reference's CPU code
        for s in slots:
            for p in s["refs"] + [s["cur"]]:
                p.free()
        for s in e2e_slots:
            for p in s["refs"] + [s["cur"]]:
                p.free()
        del pinned
        from tools import bench_extra
        try:
            extra = bench_extra.run(torch, ctx, stream, dev, peak, log=lambda m: print(m, file=sys.stderr, flush=True))
        except Exception as e:      # the headline line must not be lost to a failure in an auxiliary workload
            import traceback
            traceback.print_exc()
            extra = [{"error": repr(e)}]

    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if band else "weak", "vs_baseline": None,
            "dtype": "s16", "data": "synthetic",
            "config": common_config(args, world),
            "detail": {"frames_per_step": P * F * (1 if band else world), "distinct_picture_slots": F,
                       "gop_streams_per_gpu": n_gop, "value_one_stream_per_gpu": single_fps,
                       "overlap": "the step overlaps the launches of the GOP streams, so launches x roofline.kernel_ms (one launch timed alone on one stream) "
                                  "exceeds ms_per_step; it fits into value_one_stream_per_gpu",
                       "calls_per_picture": ("xb200_recon_frame_dev (band, stores fanned out to the peer GPUs over NVLink) + stream barrier + xb200_pad" if p2p else
                                             "xb200_recon_frame_dev (band) + NCCL all-gather of bands + xb200_pad") if band else "xb200_recon_frame_dev + xb200_pad",
                       "parallelism": (f"ctu-row bands x{world} (peer stores fused into the kernel)" if p2p else f"ctu-row bands x{world} (one all-gather per picture)") if band else f"gop-sharded x{world}",
                       "l2": f"{F} distinct picture slots x ~{(alg + 2 * w * h * 3) / 1e6:.0f} MB per rotation per GPU, {P} rotations per step; the slots share "
                             f"{len(frames)} distinct CU arrays / coefficient streams, every slot holds its own device copy"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": counters.get("traffic_bytes_per_launch"), "kernel": "k_recon_inter_v2", "kernel_ms": kern_ms, "algorithmic_bytes": alg, "peak_source": peak_src,
                         # the kernel is integer-issue bound, so the instruction rate belongs next to the bandwidth fraction.  Instructions per
                         # launch and issue utilisation come from the committed ncu capture (a profiler counter), the rate uses the time measured here
                         "integer_issue": None if not counters.get("warp_instructions_per_launch") else {
                             "warp_instructions_per_launch": counters["warp_instructions_per_launch"],
                             "lane_instructions_per_sample": round(32.0 * counters["warp_instructions_per_launch"] / (w * h * 1.5), 1),
                             "achieved_warp_inst_per_s": counters["warp_instructions_per_launch"] / (kern_ms * 1e-3),
                             "peak_warp_inst_per_s": 148 * 4 * 1.965e9, "issue_util_ncu": counters.get("issue_active_pct", 0) / 100.0,
                             "source": counters.get("file")}},
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "contexts": n_ctx, "checksum": checksum,
                    "call": "xb200_recon_frame_sparse (coefficient stream as (position, level) entries, expanded on the device) + xb200_pad + xb200_pic_download",
                    "dense_stream": {"value": e2e_dense_fps, "h2d_bytes_per_step": h2d_dense, "call": "xb200_recon_frame (dense int16 coefficient stream)"}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["extra"] = extra
        if scatter_gather is not None:
            line["nccl_scatter_gather"] = scatter_gather
        emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_OUT = None


def emit(text):
    """the ONE result line goes to the process's real stdout; everything else that lands on file descriptor 1 - e.g. the "NCCL version ..."
    banner NCCL prints from C when the communicator is created - was redirected to stderr at start-up (main)"""
    out = _RESULT_OUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k-2A", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=128, help="pictures per step and GPU (rounded down to a multiple of --slots)")
    ap.add_argument("--slots", type=int, default=16, help="distinct picture slots a step rotates through (inputs + outputs > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the non-headline workloads (tools/bench_extra.py)")
    ap.add_argument("--sharding", default="gop", choices=["gop", "band", "band-p2p"],
                    help="gop: independent pictures per GPU (weak scaling, default); band: every picture split into CTU-row bands across the GPUs, "
                         "one NCCL all-gather per picture (strong scaling, BASELINE config 4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
