import sys, time
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np, ctypes as C, homerhevc_b200 as hb
from homerhevc_b200 import synth
from homerhevc_b200.intra_jobs import presearch_jobs
w, h = 1920, 1080
tex = synth.make_texture(w, h); f = synth.make_frame(tex, w, h, 0)
t = time.perf_counter(); jobs, adi, off = presearch_jobs(f[0]); print("presearch_jobs (host prep) %.2f ms" % ((time.perf_counter() - t) * 1e3))
ctx = hb.Context(0); fr = hb.Frame(ctx, w, h); fr.upload_u8(*f); ctx.sync()
for _ in range(3): ctx.intra_presearch(fr, jobs, adi)
n = len(jobs)
t = time.perf_counter()
for _ in range(10):
    rec = np.zeros((n, 6), np.int32); rec[:, 1:4] = jobs; rec[:, 4] = -1; rec[:, 5] = -1
    arr = (hb.IntraJob * n).from_buffer(rec); sads = np.zeros((n, 35), np.uint32)
print("python-side packing %.2f ms" % ((time.perf_counter() - t) * 100))
L = ctx.L
t = time.perf_counter()
for _ in range(10):
    L.hb_intra_run(ctx.h, fr.h, None, arr, n, adi.ctypes.data_as(C.POINTER(C.c_int16)), sads.ctypes.data_as(C.POINTER(C.c_uint32)))
print("hb_intra_run C call %.2f ms" % ((time.perf_counter() - t) * 100))
for size in (32, 16, 8, 4):
    m = jobs[:, 2] == size
    rec2 = np.ascontiguousarray(rec[m]); a2 = np.concatenate([adi[o:o + 4 * size + 1] for o in off[m]]) if size >= 16 else adi[off[m][0]:off[m][0] + m.sum() * (4 * size + 1)]
    arr2 = (hb.IntraJob * len(rec2)).from_buffer(rec2); s2 = np.zeros((len(rec2), 35), np.uint32)
    L.hb_intra_run(ctx.h, fr.h, None, arr2, len(rec2), a2.ctypes.data_as(C.POINTER(C.c_int16)), s2.ctypes.data_as(C.POINTER(C.c_uint32)))
    t = time.perf_counter()
    for _ in range(5):
        L.hb_intra_run(ctx.h, fr.h, None, arr2, len(rec2), a2.ctypes.data_as(C.POINTER(C.c_int16)), s2.ctypes.data_as(C.POINTER(C.c_uint32)))
    print("size %2d: %6d jobs %.2f ms per call" % (size, len(rec2), (time.perf_counter() - t) * 200))
