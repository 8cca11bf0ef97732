"""N > 1 host logic on CPU: two gloo processes decode disjoint GOPs (the CPU oracle stands in for the device call, there is
no GPU here) and rank 0 reassembles the stream in display order; must equal the single-process result."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from xevd_b200 import dist as xdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _decode_gop(gop, n_pic=2):
    """a 'GOP' = n_pic pictures, each predicted from the previous one (a real P chain), seeded by the GOP index"""
    from oracle.pyoracle import Oracle
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    o = Oracle()
    w, h, bd = 64, 64, 10
    ref = synth.make_refs(w, h, bd, 1, seed=1000 + gop)[0]
    out = []
    for poc in range(n_pic):
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=10 * gop + poc, n_refs=1, bi_frac=0.0)
        cur = o.recon_frame(prm, HostPicture(w, h, poc), [ref], [], cl)
        o.pad(cur)
        out.append((gop, poc, hashlib.md5(cur.y.tobytes() + cur.u.tobytes() + cur.v.tobytes()).hexdigest()))
        ref = cur
    return out


def _worker(rank, world, port, n_gops, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = xdist.init("gloo")
    assert (r, w) == (rank, world)
    local = []
    for g in xdist.my_gops(n_gops, rank, world):
        local += _decode_gop(g)
    xdist.barrier()
    t = xdist.max_over_ranks(1.0 + rank)
    assert t == float(world)                      # the slowest rank (largest value) wins
    merged = xdist.gather_in_display_order(local)
    if rank == 0:
        q.put(merged)
    xdist.barrier()


def test_gop_shards_partition():
    for world in (1, 2, 3, 8):
        sh = xdist.gop_shards(13, world)
        assert sorted(g for s in sh for g in s) == list(range(13))
        assert max(len(s) for s in sh) - min(len(s) for s in sh) <= 1


@pytest.mark.timeout(300)
def test_two_ranks_gloo_equals_single_process():
    n_gops, world = 5, 2
    want = [d for g in range(n_gops) for (_, _, d) in _decode_gop(g)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_gops, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
