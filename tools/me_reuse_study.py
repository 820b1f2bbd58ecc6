"""How much of the children's search could be served from the SAD partial sums of their ancestors?  (CPU study with the oracle.)
For every CTU of a synthetic pair: run the search 64 -> 32 -> 16 -> 8 as the pre-pass does (zero predictors, parent's vector as extra
start), record every integer probe (displacement) and the (integer winner, half-pel winner) of every PU, and count the probes of the
PUs of depth >= 1 whose displacement was already probed by an ancestor (the 64x64 PU only / any ancestor).
usage: python tools/me_reuse_study.py [720p|1080p] [ctu_step]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from _frames import clip_pair, oracle_me
from _oracle import oracle
w, h = {"720p": (1280, 720), "1080p": (1920, 1080), "cif": (352, 288)}[sys.argv[1] if len(sys.argv) > 1 else "720p"]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
O = oracle()
O.orc_me_trace.argtypes = [C.c_void_p, C.c_int]; O.orc_me_trace_count.restype = C.c_int
cur, ref = clip_pair(w, h, n=2, noise=noise)
buf = np.zeros(4096, np.int16)
def search(x, y, s, starts):
    O.orc_me_trace(buf.ctypes.data, len(buf))
    r = oracle_me(cur, ref, w, h, x, y, s, 32, [(0, 0), (0, 0)], starts, 650.0)
    n = O.orc_me_trace_count(); O.orc_me_trace(None, 0)
    t = buf[:n].reshape(-1, 2)
    cut = np.where(t[:, 0] == 0x7fff)[0]
    probes = [tuple(p) for p in (t[:cut[0]] if len(cut) else t)]
    half = tuple(t[cut[0] + 1]) if len(cut) else (0, 0)
    iw = ((r.mv.x - r.subpix.x) >> 2, (r.mv.y - r.subpix.y) >> 2)
    return r, probes, iw, half
tot = {1: 0, 2: 0, 3: 0}; hit0 = dict(tot); hita = dict(tot); sub_hit = dict(tot); sub_tot = dict(tot); hitp = dict(tot)
ctus = 0
for cy in range(0, h // 64, step):
    for cx in range(0, w // 64, step):
        ctus += 1
        res = {}
        for d in range(4):
            s = 64 >> d
            for ly in range(1 << d):
                for lx in range(1 << d):
                    par = res.get((d - 1, lx // 2, ly // 2))
                    starts = []
                    if par and par[0].mv.x != 0 and par[0].mv.y != 0:
                        starts = [(par[0].mv.x, par[0].mv.y)]
                    r = search(cx * 64 + lx * s, cy * 64 + ly * s, s, starts)
                    anc = set(); a0 = set(res[(0, 0, 0)][1]) if d else set(); subs = set()
                    k = (d - 1, lx // 2, ly // 2)
                    while k[0] >= 0:
                        anc |= set(res[k][1]); subs.add((res[k][2], res[k][3])); k = (k[0] - 1, k[1] // 2, k[2] // 2)
                    res[(d, lx, ly)] = r
                    if d:
                        tot[d] += len(r[1]); hit0[d] += sum(p in a0 for p in r[1]); hita[d] += sum(p in anc for p in r[1])
                        hitp[d] += sum(p in set(par[1]) for p in r[1])
                        sub_tot[d] += 1; sub_hit[d] += ((r[2], r[3]) in subs) + 0.5 * (((r[2], r[3]) not in subs) and any(r[2] == q[0] for q in subs))
print(f"{w}x{h}, every {step}th CTU ({ctus}), noise {noise}")
for d in (1, 2, 3):
    print(f"depth {d}: {tot[d]} integer probes; already probed by the 64x64 PU {100*hit0[d]/tot[d]:.1f} %, by the parent {100*hitp[d]/tot[d]:.1f} %, by any ancestor {100*hita[d]/tot[d]:.1f} %;"
          f" sub-pel stage identical to an ancestor's (same integer + half-pel winner; half credit for same integer winner) {100*sub_hit[d]/sub_tot[d]:.1f} %")
