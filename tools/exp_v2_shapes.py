import sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
from xevd_b200 import synth
from xevd_b200.device import Context
w, h, bd = 3840, 2160, 10
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev)
ctx = Context(0); ctx.set_stream(stream.cuda_stream)
refs = synth.make_refs(w, h, bd, 2, seed=7)
drefs = [ctx.pic_alloc(w, h).upload(r) for r in refs]
curs = [ctx.pic_alloc(w, h) for _ in range(6)]
def up(cl):
    return dict(cl=cl, cus=torch.from_numpy(cl.cus.view(np.uint8).copy()).to(dev), first=torch.from_numpy(cl.ctu_first.view(np.int32).copy()).to(dev),
                ext=torch.from_numpy(cl.ext.view(np.uint8).copy()).to(dev), coef=torch.from_numpy(cl.coef.copy()).to(dev), max_cu=int(np.diff(cl.ctu_first.astype(np.int64)).max()))
def run(name, prm, wk, r0, r1):
    cl = wk["cl"]
    def fn(i): ctx.recon_frame_dev(prm, curs[i], r0, r1, wk["cus"].data_ptr(), cl.n_cu, wk["first"].data_ptr(), cl.n_ctu, wk["ext"].data_ptr(), len(cl.ext), wk["coef"].data_ptr(), cl.coef.size, has_intra=False, max_cu_per_ctu=wk["max_cu"])
    for i in range(6): fn(i)
    ctx.sync(); ts = []
    for _ in range(4):
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(i); e1.record(stream); ts.append((e0, e1))
    ctx.sync()
    print(name, round(1e3 * float(np.median([a.elapsed_time(b) for a, b in ts])), 1), "us, max_cu", wk["max_cu"])
for name, kw, nl in [("A uni 16x16", dict(variant="A"), 1), ("A bi 16x16", dict(variant="A", bi_frac=1.0), 2), ("B uni quadtree", dict(variant="B", bi_frac=0.0), 1),
                     ("B 50% bi quadtree", dict(variant="B"), 2), ("A uni 8x8", dict(variant="A", log2_cu=3), 1), ("A uni 32x32", dict(variant="A", log2_cu=5), 1), ("A uni 64x64", dict(variant="A", log2_cu=6), 1)]:
    prm, cl = synth.make_inter_frame(w, h, bit_depth=bd, seed=1, n_refs=2 if nl == 2 else 1, **kw)
    run(name, prm, up(cl), drefs[:1] if nl == 1 else drefs, [] if nl == 1 else drefs[::-1])
