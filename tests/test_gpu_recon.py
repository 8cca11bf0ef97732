"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs."""
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


def _run(ctx, oracle, w, h, bd, variant, seed, n_refs=2, **kw):
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, variant=variant, seed=seed, n_refs=n_refs, **kw)
    cl.validate()
    refs = synth.make_refs(w, h, bd, n_refs, seed=seed + 100)
    want = oracle.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
    drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
    cur = ctx.pic_alloc(w, h)
    ctx.recon_frame(prm, cur, drefs, drefs[::-1], cl)
    got = cur.download(maps=True)
    for p in drefs + [cur]:
        p.free()
    for a, b, n in zip(got.planes(), want.planes(), "YUV"):
        assert np.array_equal(a, b), f"plane {n}: {int((a != b).sum())} samples differ"
    assert np.array_equal(got.map_mv, want.map_mv)
    assert np.array_equal(got.map_refi, want.map_refi)
    assert np.array_equal(got.map_scu, want.map_scu)


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("bd", [8, 10])
def test_inter_small(ctx, oracle, variant, bd):
    _run(ctx, oracle, 256, 136, bd, variant, seed=3, coded_frac=0.8)


def test_inter_large_mv_clipping(ctx, oracle):
    # vectors far outside the picture: exercises xevd_mv_clip and the T3 variant/phase split
    _run(ctx, oracle, 192, 128, 10, "B", seed=4, mv_range_px=400)


def test_inter_iqt(ctx, oracle):
    _run(ctx, oracle, 256, 128, 10, "B", seed=5, iqt=True)


def test_inter_main_tables(ctx, oracle):
    _run(ctx, oracle, 256, 128, 10, "B", seed=6, main_mv=True)


def test_inter_8x8_cus(ctx, oracle):
    _run(ctx, oracle, 128, 64, 10, "A", seed=7, log2_cu=3)


def test_inter_1080p(ctx, oracle):
    _run(ctx, oracle, 1920, 1080, 10, "A", seed=8, n_refs=1)


def test_pad(ctx, oracle):
    rng = np.random.default_rng(1)
    p = HostPicture.random(200, 104, 10, rng)
    d = ctx.pic_alloc(200, 104).upload(p, padded=True)
    got = d.download_padded()
    want = oracle.pad(p.copy())
    assert np.array_equal(got.buf_y, want.buf_y) and np.array_equal(got.buf_u, want.buf_u) and np.array_equal(got.buf_v, want.buf_v)
    d.free()
