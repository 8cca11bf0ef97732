import sys, os; sys.path.insert(0,'.')
import numpy as np
from xevd_b200 import synth
from xevd_b200.device import Context
from xevd_b200.frame import HostPicture
from oracle.pyoracle import Oracle
o=Oracle()
w,h,bd=64,64,10
c=Context(0)
prm, cl = synth.make_inter_frame(w,h,bit_depth=bd,variant="A",seed=9,n_refs=1,log2_cu=2)
cl.cus["mv"][:]=0
refs = synth.make_refs(w,h,bd,1,seed=109)
for pl in refs[0].planes(): pl[:]=512
refs[0].pad_borders()
want=o.recon_frame(prm,HostPicture(w,h,prm.poc),refs,refs[::-1],cl)
drefs=[c.pic_alloc(w,h).upload(r) for r in refs]
cur=c.pic_alloc(w,h)
c.recon_frame(prm,cur,drefs,drefs[::-1],cl)
got=cur.download()
for a,b,n in zip(got.planes(),want.planes(),"YUV"):
    bad=np.argwhere(a!=b)
    print(n,len(bad))
print("coef CU0..3", cl.coef[:128].reshape(4,32))
print("got V\n", got.v[:4,:8]-512, "\nwant V\n", want.v[:4,:8]-512)
print("got U\n", got.u[:4,:8]-512, "\nwant U\n", want.u[:4,:8]-512)
