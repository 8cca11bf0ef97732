import sys, os; sys.path.insert(0,'.')
import numpy as np
from xevd_b200 import synth
from xevd_b200.device import Context
from xevd_b200.frame import HostPicture
from oracle.pyoracle import Oracle
o=Oracle()
w,h=256,136
kw,bd=dict(log2_ctu=5),10
for mode in ("full","nocoef","zeromv","only105","drop_others_in_ctu"):
    prm, cl = synth.make_inter_frame(w,h,bit_depth=bd,variant="C",seed=31,n_refs=2,coded_frac=0.8,**kw)
    if mode=="nocoef": cl.cus["cbf"]=0
    if mode=="zeromv": cl.cus["mv"]=0
    refs=synth.make_refs(w,h,bd,2,seed=131)
    want=o.recon_frame(prm,HostPicture(w,h,prm.poc),refs,refs[::-1],cl)
    c=Context(0)
    drefs=[c.pic_alloc(w,h).upload(r) for r in refs]
    cur=c.pic_alloc(w,h)
    c.recon_frame(prm,cur,drefs,drefs[::-1],cl)
    got=cur.download(maps=True)
    print(mode,[int((a!=b).sum()) for a,b in zip(got.planes(),want.planes())])
    if mode=="full":
        ctu=(80//32)*(w//32)+(24//32)
        a,b=int(cl.ctu_first[ctu]),int(cl.ctu_first[ctu+1])
        print("CTU",ctu,"cus",a,b)
        for i in range(a,b): print(i, cl.cus[i])
    c.close()
