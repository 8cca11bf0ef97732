import sys, os; sys.path.insert(0,'.')
import numpy as np
from xevd_b200 import synth
from xevd_b200.device import Context
from xevd_b200.frame import HostPicture
from oracle.pyoracle import Oracle
o=Oracle()
def smooth(pic,bd):
    for pl in pic.planes(): pl[...] = (pl.astype(np.int32)//8 + (1<<(bd-1))).astype(np.int16)
w,h,bd=320,200,10
prm, cl = synth.make_inter_frame(w,h,bit_depth=bd,variant="B",seed=51,n_refs=2,coded_frac=0.6,mv_range_px=24)
cl.cus["qp_map"]=np.random.default_rng(3).integers(30,46,cl.n_cu)
refs=synth.make_refs(w,h,bd,2,seed=52)
for r in refs: smooth(r,bd); r.pad_borders()
tbl=synth.chroma_qp_table(False)
want=o.recon_frame(prm,HostPicture(w,h,prm.poc),refs,refs[::-1],cl)
c=Context(0)
drefs=[c.pic_alloc(w,h).upload(r) for r in refs]
cur=c.pic_alloc(w,h)
c.recon_frame(prm,cur,drefs,drefs[::-1],cl)
got=cur.download(maps=True)
print("recon", [int((a!=b).sum()) for a,b in zip(got.planes(),want.planes())], "maps", np.array_equal(got.map_scu,want.map_scu), np.array_equal(got.map_mv,want.map_mv), np.array_equal(got.map_refi,want.map_refi), "edge", np.array_equal(cur.download_edge_map(), cl.edge_flags()))
for a,b,n in zip(got.planes(),want.planes(),"YUV"):
    bad=np.argwhere(a!=b)
    if len(bad): print("recon",n,len(bad),bad.min(0),bad.max(0))
idx=[i for i,cu in enumerate(cl.cus) if cu["x"]<=bad[:,1].min()*2+1 < cu["x"]+(1<<cu["log2w"]) and cu["y"]<=bad[:,0].min()*2+1 < cu["y"]+(1<<cu["log2h"])]
print(cl.cus[idx])
import os
os.environ["XB200_FORCE_GENERIC"]="1"
c2=Context(0)
dr2=[c2.pic_alloc(w,h).upload(r) for r in refs]
cur2=c2.pic_alloc(w,h)
c2.recon_frame(prm,cur2,dr2,dr2[::-1],cl)
g2=cur2.download()
print("generic kernel recon", [int((a!=b).sum()) for a,b in zip(g2.planes(),want.planes())])
o.deblock_frame(prm,want,cl,tbl)
c.deblock(prm,cur)
got=cur.download()
for a,b,n in zip(got.planes(),want.planes(),"YUV"):
    bad=np.argwhere(a!=b); print("deblock",n,len(bad),bad[:8].tolist())
