import sys, time
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, homerhevc_b200 as hb
from homerhevc_b200 import synth
w,h=1920,1080
tex=synth.make_texture(w,h)
c=hb.Context(0); pp=hb.Prepass(c,w,h,qp=32,use_graph=1)
fb=w*h*3//2
pin=c.pinned(fb*5); host=[]
for n in range(5):
    y,u,v=synth.make_frame(tex,w,h,n); b=pin[n*fb:(n+1)*fb]
    py=b[:w*h].reshape(h,w); pu=b[w*h:w*h*5//4].reshape(h//2,w//2); pv=b[w*h*5//4:].reshape(h//2,w//2); py[:],pu[:],pv[:]=y,u,v; host.append((py,pu,pv))
cur,ref=hb.Frame(c,w,h),hb.Frame(c,w,h)
n=pp.num_ctus(); tables=c.pinned(pp.tables_bytes()); out=c.pinned(fb+4*w*h); sel=np.zeros(n,np.uint8); off=np.zeros(n+1,np.int32)
T={k:0.0 for k in ('upload','run','fetch_tables','sync1','select','gather','sync2')}
def tick(k,t0): T[k]+=time.perf_counter()-t0
N=50
for i in range(N+5):
    if i==5: T={k:0.0 for k in T}
    t=time.perf_counter(); cur.upload_u8(*host[i%4+1]); ref.upload_u8(*host[i%4]); tick('upload',t)
    t=time.perf_counter(); pp.run(cur,ref,650.0); tick('run',t)
    t=time.perf_counter(); pp.fetch_tables(tables); tick('fetch_tables',t)
    t=time.perf_counter(); c.sync(); tick('sync1',t)
    t=time.perf_counter(); pp.select(tables,60,sel,off); tick('select',t)
    t=time.perf_counter(); pp.gather(sel,off,out); tick('gather',t)
    t=time.perf_counter(); c.sync(); tick('sync2',t)
print({k:round(v/N*1e3,4) for k,v in T.items()}, 'ms per frame; total', round(sum(T.values())/N*1e3,3))
