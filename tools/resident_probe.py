"""Stage times of the device-resident per-frame flow on ONE stream (each stage followed by a wait): where a frame's latency goes."""
import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, homerhevc_b200 as hb
from homerhevc_b200 import synth
from homerhevc_b200.lib import SAO_PARAM_DT, sao_decide_from_candidates
w, h = 1920, 1080
tex = synth.make_texture(w, h)
ctx = hb.Context(0)
fb = w * h * 3 // 2
pin = ctx.pinned(fb * 6)
planes = []
for i in range(6):
    y, u, v = synth.make_frame(tex, w, h, i)
    b = pin[i * fb:(i + 1) * fb]
    py = b[:w * h].reshape(h, w); pu = b[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); pv = b[w * h * 5 // 4:].reshape(h // 2, w // 2)
    py[:], pu[:], pv[:] = y, u, v
    planes.append((py, pu, pv))
pp = hb.Prepass(ctx, w, h, qp=32, use_graph=1, compact_tables=2)
n = pp.num_ctus()
cur, rec, refs = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h), [hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)]
refs[0].upload_u8(*planes[0]); ctx.sync()
tables = ctx.pinned(pp.tables_bytes()); levels = ctx.pinned(4 * w * h)
sel = np.zeros(n, np.uint8); off = np.zeros(n + 1, np.int32); prm = np.zeros(n, SAO_PARAM_DT)
lam = (60.0, 48.0, 48.0)
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
which = 0
for it in range(12):
    j = 1 + it % 5
    t = time.perf_counter(); cur.upload_u8(*planes[j]); ctx.sync(); tick("upload cur", t)
    t = time.perf_counter(); pp.run(cur, refs[which], 650.0); ctx.sync(); tick("pre-pass", t)
    t = time.perf_counter(); pp.fetch_tables(tables); ctx.sync(); tick("fetch tables", t)
    t = time.perf_counter(); pp.select(tables, 60, sel, off); tick("select (host)", t)
    t = time.perf_counter(); pp.finalise(sel, off, rec, levels, 2, 2); ctx.sync(); tick("finalise: gather+units+deblock+pad+levels d2h", t)
    t = time.perf_counter(); cand = ctx.sao_candidates(cur, rec, lam); tick("sao stats+derive+d2h", t)
    t = time.perf_counter(); p = sao_decide_from_candidates(cand, lam); tick("sao decide (host)", t)
    t = time.perf_counter(); ctx.sao_apply(rec, refs[1 - which], p["type"], p["offset"]); tick("sao apply+pad", t)
    which = 1 - which
    t = time.perf_counter()
    pp.frame_begin_resident(cur, refs[which], planes[1 + (it + 1) % 5], 650.0, tables)
    nlev = pp.frame_finish_resident(cur, 60, tables, sel, off, rec, refs[1 - which], (2, 2, 0, 0), lam, levels, prm); ctx.sync()
    which = 1 - which
    tick("begin+finish resident (one call pair)", t)
for k, v in T.items():
    print("%-50s %.3f ms" % (k, float(np.median(v[3:]))))
print("levels bytes", nlev, "tables bytes", tables.nbytes)
