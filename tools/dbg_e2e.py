import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, homerhevc_b200 as hb
from homerhevc_b200 import synth
w,h=int(sys.argv[1]),int(sys.argv[2]); nslots=int(sys.argv[3])
tex=synth.make_texture(w,h)
host=[synth.make_frame(tex,w,h,n) for n in range(5)]
slots=[]
for k in range(nslots):
    c=hb.Context(0); pp=hb.Prepass(c,w,h,qp=32,use_graph=int(sys.argv[4]))
    n=pp.num_ctus()
    slots.append(dict(ctx=c,cur=hb.Frame(c,w,h),ref=hb.Frame(c,w,h),pp=pp,tables=c.pinned(pp.tables_bytes()),out=c.pinned(w*h*3//2+4*w*h),sel=np.zeros(n,np.uint8),off=np.zeros(n+1,np.int32)))
for i in range(8):
    sl=slots[i%nslots]
    sl['cur'].upload_u8(*host[i%4+1]); sl['ref'].upload_u8(*host[i%4])
    sl['pp'].run(sl['cur'],sl['ref'],650.0)
    sl['pp'].fetch_tables(sl['tables'])
    sl['ctx'].sync(); print('step',i,'run ok',flush=True)
    sl['pp'].select(sl['tables'],60,sl['sel'],sl['off'])
    print(' sel hist',np.bincount(sl['sel'],minlength=5),'levels',sl['off'][-1],flush=True)
    nb=sl['pp'].gather(sl['sel'],sl['off'],sl['out'])
    sl['ctx'].sync(); print(' gather ok',nb,flush=True)
