"""Which stage bounds bench.py's e2e?  The threaded per-stream loop with stages switched off one at a time."""
import sys, time, threading
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np, homerhevc_b200 as hb
from homerhevc_b200 import synth
w, h = 1920, 1080
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tex = synth.make_texture(w, h)
c0 = hb.Context(0)
fb = w * h * 3 // 2
pin = c0.pinned(fb * 5); host = []
for n in range(5):
    y, u, v = synth.make_frame(tex, w, h, n); b = pin[n * fb:(n + 1) * fb]
    py = b[:w * h].reshape(h, w); pu = b[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); pv = b[w * h * 5 // 4:].reshape(h // 2, w // 2); py[:], pu[:], pv[:] = y, u, v; host.append((py, pu, pv))
slots = []
for k in range(S):
    c = hb.Context(0); pp = hb.Prepass(c, w, h, qp=32, use_graph=1); n = pp.num_ctus()
    sl = dict(c=c, pp=pp, cur=hb.Frame(c, w, h), ref=hb.Frame(c, w, h), tables=c.pinned(pp.tables_bytes()), out=c.pinned(fb + 4 * w * h),
              sel=np.zeros(n, np.uint8), off=np.zeros(n + 1, np.int32))
    sl['cur'].upload_u8(*host[1]); sl['ref'].upload_u8(*host[0]); c.sync(); slots.append(sl)
def frame(sl, i, up, run, tab, sel, gat):
    j = i % 4
    if up: sl['cur'].upload_u8(*host[j + 1]); sl['ref'].upload_u8(*host[j])
    if run: sl['pp'].run(sl['cur'], sl['ref'], 650.0)
    if tab: sl['pp'].fetch_tables(sl['tables'])
    sl['c'].sync()
    if sel: sl['pp'].select(sl['tables'], 60, sl['sel'], sl['off'])
    if gat: sl['pp'].gather(sl['sel'], sl['off'], sl['out']); sl['c'].sync()
def bench(name, n=160, **kw):
    def worker(k):
        for i in range(k, n, S): frame(slots[k], i, **kw)
    for rep in range(2):
        ths = [threading.Thread(target=worker, args=(k,)) for k in range(S)]
        t = time.perf_counter()
        for x in ths: x.start()
        for x in ths: x.join()
        dt = time.perf_counter() - t
    print(f"{name:40s} {n / dt:8.0f} fps")
full = dict(up=1, run=1, tab=1, sel=1, gat=1)
bench("full", **full)
bench("no upload", **{**full, 'up': 0})
bench("no gather", **{**full, 'gat': 0})
bench("no select/gather", **{**full, 'sel': 0, 'gat': 0})
bench("no upload, no gather", **{**full, 'up': 0, 'gat': 0})
bench("run + sync only", up=0, run=1, tab=0, sel=0, gat=0)
bench("upload + sync only", up=1, run=0, tab=0, sel=0, gat=0)
bench("upload + run", up=1, run=1, tab=0, sel=0, gat=0)
bench("tables + select + gather (no run)", up=0, run=0, tab=1, sel=1, gat=1)
bench("gather only", up=0, run=0, tab=0, sel=0, gat=1)
