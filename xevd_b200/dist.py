"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous / barrier / reductions.

Frames of one EVC stream shard across GPUs only at closed-GOP (IDR) boundaries: an IDR flushes the DPB
(picman_flush_pb, src_base/xevd_picman.c:112-156) and restarts POC, so GOPs are independent units.  Ranks therefore
decode disjoint GOPs with no data-path collective; the only communication is the barrier around a timed region, the
max-over-ranks of the elapsed time, and (for a consumer that wants one ordered stream) gathering decoded pictures or
their digests in display order on rank 0.
"""
from __future__ import annotations

import os
from typing import Any, List, Sequence


import contextlib


def _null():
    return contextlib.nullcontext()


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gop_shards(n_gops: int, world: int) -> List[List[int]]:
    """round-robin assignment of GOP indices to ranks (keeps every rank busy from the first GOP on)"""
    return [list(range(r, n_gops, world)) for r in range(world)]


def my_gops(n_gops: int, rank: int, world: int) -> List[int]:
    return gop_shards(n_gops, world)[rank]


def band_partition(h: int, log2_ctu: int, world: int) -> List[tuple]:
    """(first CTU row, CTU rows) per rank: contiguous bands of CTU rows, sizes differing by at most one row (8K has
    68 CTU rows -> four bands of 9 and four of 8)"""
    n = (h + (1 << log2_ctu) - 1) >> log2_ctu
    base, extra = divmod(n, world)
    out, r0 = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append((r0, k))
        r0 += k
    return out


class BandExchange:
    """Per-picture exchange step of intra-picture sharding: every rank has reconstructed its band of the current picture into its
    own copy of the device picture; one in-place all-gather over NCCL (NVLink / NVSwitch) of the packed bands - planes plus
    per-SCU maps - completes the picture on every GPU, where it serves as a reference for the next pictures.
    Chunks are sized for the largest band so that a single equal-sized all-gather does the job."""

    def __init__(self, ctx, pic, log2_ctu: int, rank: int, world: int, device):
        import torch
        self.ctx, self.rank, self.world = ctx, rank, world
        self.bands = band_partition(pic.h, log2_ctu, world)
        ctu = 1 << log2_ctu
        self.rows = [(r0 * ctu, min(k * ctu, pic.h - r0 * ctu)) for r0, k in self.bands]       # luma rows (y0, rows)
        self.chunk = max(ctx.band_bytes(pic, rows) for _, rows in self.rows if rows > 0)
        self.buf = torch.empty(self.chunk * world, dtype=torch.uint8, device=device)

    def exchange(self, pic):
        """pack -> all-gather -> unpack, all in the order of the CONTEXT's stream: the collective is issued with that stream as torch's
        current stream, whatever the caller's current stream is (pack / unpack are launched on the context stream by the library)"""
        import torch
        import torch.distributed as dist
        base = self.buf.data_ptr()
        y0, rows = self.rows[self.rank]
        if rows > 0:
            self.ctx.band_pack(pic, y0, rows, base + self.rank * self.chunk)
        if self.world > 1:
            mine = self.buf[self.rank * self.chunk:(self.rank + 1) * self.chunk]
            st = self.ctx.stream
            with torch.cuda.stream(torch.cuda.ExternalStream(st)) if (st and self.buf.is_cuda) else _null():
                dist.all_gather_into_tensor(self.buf, mine)             # in place: rank r's chunk is slot r of the output
        for r, (yr, nr) in enumerate(self.rows):
            if r != self.rank and nr > 0:
                self.ctx.band_unpack(pic, yr, nr, base + r * self.chunk)


def open_peer_pictures(ctx, pic):
    """band mode with peer stores: exchange the CUDA IPC handles of every rank's copy of `pic` (host plumbing, once per picture
    buffer) and map the other ranks' copies into this process"""
    import ctypes as C
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    h = (C.c_ubyte * 64)()
    ctx._chk(ctx.lib.xb200_pic_export(ctx.handle, pic.handle, h), "xb200_pic_export")
    handles = [None] * dist.get_world_size()
    dist.all_gather_object(handles, bytes(h))
    for r, hb in enumerate(handles):
        if r != dist.get_rank():
            buf = (C.c_ubyte * 64).from_buffer_copy(hb)
            ctx._chk(ctx.lib.xb200_pic_open_peer(ctx.handle, pic.handle, buf), "xb200_pic_open_peer")


def init(backend: str | None = None, device_index: int | None = None):
    """initialise torch.distributed from the torchrun environment (no-op for a single process)"""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world == 1 or dist.is_initialized():
        return rank, world, local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local if device_index is None else device_index)
        kw["device_id"] = torch.device("cuda", local if device_index is None else device_index)
    dist.init_process_group(backend, **kw)
    return rank, world, local


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """the slowest rank defines the time of a step"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_display_order(local: Sequence[tuple], dst: int = 0) -> List[Any] | None:
    """local: [(gop_index, poc, payload), ...] of this rank.  Returns, on rank dst, every rank's payloads sorted by
    (gop_index, poc) - the order xevd_pull would have delivered them from a single decoder - and None elsewhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [p for _, _, p in sorted(local, key=lambda t: (t[0], t[1]))]
    rank, world = dist.get_rank(), dist.get_world_size()
    bucket: List[Any] = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst)
    if rank != dst:
        return None
    merged = [item for part in bucket for item in part]
    return [p for _, _, p in sorted(merged, key=lambda t: (t[0], t[1]))]
