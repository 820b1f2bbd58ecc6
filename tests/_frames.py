"""helpers shared by the GPU parity tests: synthetic frames, padded int16 views for the oracle"""
import ctypes as C

import numpy as np

import homerhevc_b200 as hb
from homerhevc_b200 import synth
from _oracle import OrcMeIn, OrcMeOut, OrcMv, OrcTuOut, oracle, ptr

PAD = 96


def padded_i16(plane_u8, pad=PAD):
    return np.ascontiguousarray(np.pad(plane_u8.astype(np.int16), pad, mode="edge"))


class HostFrame:
    """u8 planes + border-replicated int16 copies the oracle can address like the reference's wnd_t"""

    def __init__(self, y, u, v):
        self.y, self.u, self.v = y, u, v
        self.p = [padded_i16(y, PAD), padded_i16(u, PAD // 2), padded_i16(v, PAD // 2)]

    def ptr(self, comp, x, y):
        pad = PAD if comp == 0 else PAD // 2
        a = self.p[comp]
        return ptr(a.reshape(-1), (pad + y) * a.shape[1] + pad + x), a.shape[1]

    def block(self, comp, x, y, n):
        pl = (self.y, self.u, self.v)[comp]
        return pl[y:y + n, x:x + n]


def clip_pair(w, h, n=1, noise=3.0, seed=None):
    tex = synth.make_texture(w, h, seed)
    cur = HostFrame(*synth.make_frame(tex, w, h, n, noise))
    ref = HostFrame(*synth.make_frame(tex, w, h, n - 1, noise))
    return cur, ref


def upload(ctx, hf, w, h):
    f = hb.Frame(ctx, w, h)
    f.upload_u8(hf.y, hf.u, hf.v)
    return f


def oracle_me(cur, ref, w, h, x, y, size, qp, amvp, starts, avg_dist, action=7):
    O = oracle()
    mi = OrcMeIn()
    mi.orig, mi.orig_stride = cur.ptr(0, x, y)
    mi.ref, mi.ref_stride = ref.ptr(0, x, y)
    mi.gx, mi.gy, mi.size, mi.frame_w, mi.frame_h, mi.range_x, mi.range_y = x, y, size, w, h, 128, 64
    mi.n_amvp = len(amvp)
    for i, (ax, ay) in enumerate(amvp):
        mi.amvp[i].x, mi.amvp[i].y = ax, ay
    mi.n_start = len(starts)
    for i, (sx, sy) in enumerate(starts):
        mi.start[i].x, mi.start[i].y = sx, sy
    mi.qp, mi.avg_dist, mi.action = qp, avg_dist, action
    mo = OrcMeOut()
    O.orc_motion_estimation(C.byref(mi), C.byref(mo))
    return mo


def oracle_mc(ref, comp, x, y, n, mvx, mvy):
    O = oracle()
    out = np.zeros((n, n), np.int16)
    p, s = ref.ptr(comp, x, y)
    (O.orc_mc_luma if comp == 0 else O.orc_mc_chroma)(p, s, ptr(out.reshape(-1)), n, n, OrcMv(mvx, mvy))
    return out.astype(np.uint8)


def oracle_mc_bi(ref0, ref1, comp, x, y, n, mv0, mv1):
    """bi-prediction: the 14-bit predictions of the two lists, averaged (hmr_motion_inter.c:3047-3056)"""
    O = oracle()
    a = np.zeros((n, n), np.int16); b = np.zeros((n, n), np.int16); out = np.zeros((n, n), np.int16)
    f = O.orc_mc_luma_ex if comp == 0 else O.orc_mc_chroma_ex
    p, s = ref0.ptr(comp, x, y); f(p, s, ptr(a.reshape(-1)), n, n, OrcMv(*mv0), 1)
    p, s = ref1.ptr(comp, x, y); f(p, s, ptr(b.reshape(-1)), n, n, OrcMv(*mv1), 1)
    O.orc_weighted_average(ptr(a.reshape(-1)), n, ptr(b.reshape(-1)), n, ptr(out.reshape(-1)), n, n, n)
    return out.astype(np.uint8)


def oracle_tu(cur_block, pred_block, n, comp, qp_eff, isl, sh, avg_dist, weight):
    O = oracle()
    o = np.ascontiguousarray(cur_block.astype(np.int16)).reshape(-1)
    p = np.ascontiguousarray(pred_block.astype(np.int16)).reshape(-1)
    co = np.zeros(n * n, np.int16); de = np.zeros(n * n, np.int16)
    to = OrcTuOut()
    O.orc_encode_inter_tu(O.tables, ptr(o), n, ptr(p), n, ptr(co), ptr(de), n, n, comp, qp_eff, isl, sh, avg_dist, weight, C.byref(to))
    return co.reshape(n, n), de.reshape(n, n).astype(np.uint8), to


def oracle_intra_tu(cur_block, pred_block, n, comp, qp_eff, scan_mode, isl, sh, weight):
    O = oracle()
    o = np.ascontiguousarray(cur_block.astype(np.int16)).reshape(-1)
    p = np.ascontiguousarray(pred_block.astype(np.int16)).reshape(-1)
    co = np.zeros(n * n, np.int16); de = np.zeros(n * n, np.int16)
    to = OrcTuOut()
    O.orc_encode_intra_tu(O.tables, ptr(o), n, ptr(p), n, ptr(co), ptr(de), n, n, comp, qp_eff, scan_mode, isl, sh, weight, C.byref(to))
    return co.reshape(n, n), de.reshape(n, n).astype(np.uint8), to
