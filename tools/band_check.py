#!/usr/bin/env python3
"""Parity check of intra-picture band sharding (BASELINE config 4) under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/band_check.py
Every rank reconstructs its CTU-row band of a chain of P pictures, the bands are all-gathered over NCCL, and the completed
pictures (planes, maps, padding) must equal a whole-picture reconstruction done on the same GPU."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    from xevd_b200 import dist as xdist, synth
    from xevd_b200.device import Context
    p2p = "--p2p" in sys.argv        # exchange fused into the kernel's stores (NVLink peer writes) instead of the all-gather
    rank, world, local = xdist.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    w, h, bd, lg = 1920, 1080, 10, 6
    ctx, ctx1 = Context(local), Context(local)
    ctx.set_stream(stream.cuda_stream)
    ref = synth.make_refs(w, h, bd, 1, seed=5)[0]
    d_ref = ctx.pic_alloc(w, h).upload(ref)
    d_ref1 = ctx1.pic_alloc(w, h).upload(ref)
    bands = xdist.band_partition(h, lg, world)
    ok = True
    with torch.cuda.stream(stream):
        ex = None
        for poc in range(3):
            prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B" if poc else "A", seed=60 + poc, n_refs=1, bi_frac=0.0)
            cur, cur1 = ctx.pic_alloc(w, h), ctx1.pic_alloc(w, h)
            ex = ex or xdist.BandExchange(ctx, cur, lg, rank, world, dev)
            r0, k = bands[rank]
            prm.ctu_row0, prm.ctu_rows = r0, k
            if p2p:
                ctx.sync()                                   # the zero-fill of the fresh pictures must not race the peers' stores
                xdist.open_peer_pictures(ctx, cur)
                xdist.barrier()
            if k > 0:
                ctx.recon_frame(prm, cur, [d_ref], [], cl.band(r0, k))
            if p2p:
                flag = torch.zeros(1, device=dev)
                if world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(flag)                    # barrier on the stream: every rank's kernel (and its peer stores) is complete
            else:
                ex.exchange(cur)
            ctx.pad(cur)
            ctx.sync()
            prm.ctu_row0 = prm.ctu_rows = 0
            ctx1.recon_frame(prm, cur1, [d_ref1], [], cl)
            ctx1.pad(cur1)
            a, b = cur.download_padded(), cur1.download_padded()
            am, bm = cur.download(maps=True), cur1.download(maps=True)
            same = all(np.array_equal(x, y) for x, y in ((a.buf_y, b.buf_y), (a.buf_u, b.buf_u), (a.buf_v, b.buf_v), (am.map_mv, bm.map_mv),
                                                          (am.map_scu, bm.map_scu), (am.map_refi, bm.map_refi)))
            print(f"[rank {rank}] picture {poc}: band {bands[rank]} {'OK' if same else 'MISMATCH'}", flush=True)
            ok &= same
            d_ref, d_ref1 = cur, cur1
    t = torch.tensor([0 if ok else 1], device=dev)
    import torch.distributed as dist
    if world > 1:
        dist.all_reduce(t)
        dist.barrier()
        dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print(f"band sharding parity OK on {world} GPU(s)" + (" [peer stores]" if p2p else " [all-gather]"))


if __name__ == "__main__":
    main()
