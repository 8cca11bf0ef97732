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


# ---- intra-picture band sharding (BASELINE config 4): host logic on CPU ------------------------------------------------------
def _band_worker(rank, world, port, q):
    """each rank reconstructs only its CTU-row band of two chained pictures (the CPU oracle stands in for the device call), the bands
    are all-gathered so that every rank owns the complete picture before it becomes the next picture's reference"""
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    xdist.init("gloo")
    o = Oracle()
    w, h, bd, lg = 192, 200, 10, 6
    bands = xdist.band_partition(h, lg, world)
    ref = synth.make_refs(w, h, bd, 1, seed=5)[0]
    digests = []
    for poc in range(2):
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=40 + poc, n_refs=1, bi_frac=0.0)
        cur = HostPicture(w, h, poc)
        r0, k = bands[rank]
        o.recon_frame(prm, cur, [ref], [], cl.band(r0, k))
        # exchange: equal-sized chunks (largest band), luma + chroma rows of the band
        rows = [(a << lg, min(b << lg, h - (a << lg))) for a, b in bands]
        chunk = max(n for _, n in rows) * w * 2
        mine = torch.zeros(2 * chunk, dtype=torch.uint8)            # gloo has no 16-bit integer all-gather: ship bytes
        y0, n = rows[rank]
        packed = np.concatenate([cur.y[y0:y0 + n].ravel(), cur.u[y0 // 2:(y0 + n) // 2].ravel(), cur.v[y0 // 2:(y0 + n) // 2].ravel()])
        mine[:2 * packed.size] = torch.from_numpy(packed.copy().view(np.uint8))
        got = [torch.zeros(2 * chunk, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(got, mine)
        for r, (yr, nr) in enumerate(rows):
            if r == rank:
                continue
            a = got[r].numpy().view(np.int16)
            cur.y[yr:yr + nr] = a[:nr * w].reshape(nr, w)
            cur.u[yr // 2:(yr + nr) // 2] = a[nr * w:nr * w + nr * w // 4].reshape(nr // 2, w // 2)
            cur.v[yr // 2:(yr + nr) // 2] = a[nr * w + nr * w // 4:nr * w + nr * w // 2].reshape(nr // 2, w // 2)
        o.pad(cur)
        digests.append(hashlib.md5(cur.buf_y.tobytes() + cur.buf_u.tobytes() + cur.buf_v.tobytes()).hexdigest())
        ref = cur
    q.put((rank, digests))
    xdist.barrier()


def test_band_partition_covers_all_rows():
    for h, lg, world in ((4320, 6, 8), (2160, 6, 8), (136, 6, 2), (1080, 7, 4), (64, 6, 8)):
        b = xdist.band_partition(h, lg, world)
        n = (h + (1 << lg) - 1) >> lg
        assert b[0][0] == 0 and sum(k for _, k in b) == n and all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(k for _, k in b) - min(k for _, k in b) <= 1


def test_cu_list_band_slices():
    from xevd_b200 import synth
    prm, cl = synth.make_inter_frame(256, 200, variant="C", seed=3, n_refs=2)
    parts = [cl.band(r0, k) for r0, k in xdist.band_partition(200, 6, 3)]
    assert sum(len(p.cus) for p in parts) == len(cl.cus) and sum(p.coef.size for p in parts) == cl.coef.size
    for p in parts:
        assert p.ctu_first[0] == 0 and p.ctu_first[-1] == len(p.cus)
        if len(p.cus):
            assert p.cus["coef_off"][0] == 0


@pytest.mark.timeout(300)
def test_two_ranks_band_sharding_equals_single_process():
    import hashlib as _h
    from oracle.pyoracle import Oracle
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    o = Oracle()
    w, h, bd = 192, 200, 10
    ref = synth.make_refs(w, h, bd, 1, seed=5)[0]
    want = []
    for poc in range(2):
        prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant="B", seed=40 + poc, n_refs=1, bi_frac=0.0)
        cur = o.recon_frame(prm, HostPicture(w, h, poc), [ref], [], cl)
        o.pad(cur)
        want.append(_h.md5(cur.buf_y.tobytes() + cur.buf_u.tobytes() + cur.buf_v.tobytes()).hexdigest())
        ref = cur
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == want and res[1] == want
