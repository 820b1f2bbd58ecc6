"""ctypes access to the checker: oracle/liboracle.so (C restatement) and, when it was built, the compiled
UNMODIFIED reference under oracle/_ref/ (libhomer_ref.so + librefdrv.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


def ptr(a, off=0):
    """int16 pointer `off` elements into a C-contiguous int16 array."""
    assert a.dtype == np.int16 and a.flags["C_CONTIGUOUS"]
    return C.cast(a.ctypes.data + 2 * off, i16p)


def aligned_i16(n, align=64):
    raw = np.zeros(n + align, dtype=np.int16)
    off = (-raw.ctypes.data % align) // 2
    return raw[off:off + n]


def build_oracle():
    if not os.path.exists(os.path.join(ORACLE_DIR, "liboracle.so")) or (
            os.path.isdir("/root/reference") and not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "librefdrv.so"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


class OrcMv(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class OrcMeIn(C.Structure):
    _fields_ = [("orig", i16p), ("orig_stride", C.c_int), ("ref", i16p), ("ref_stride", C.c_int),
                ("gx", C.c_int), ("gy", C.c_int), ("size", C.c_int), ("frame_w", C.c_int), ("frame_h", C.c_int),
                ("range_x", C.c_int), ("range_y", C.c_int),
                ("n_amvp", C.c_int), ("amvp", OrcMv * 2), ("n_start", C.c_int), ("start", OrcMv * 3),
                ("qp", C.c_int), ("avg_dist", C.c_double), ("action", C.c_int)]


class OrcMeOut(C.Structure):
    _fields_ = [("mv", OrcMv), ("subpix", OrcMv), ("sad", C.c_uint32), ("n_int_sads", C.c_uint32)]


class OrcTuOut(C.Structure):
    _fields_ = [("sum", C.c_int32), ("ssd", C.c_uint32), ("ssd_zero", C.c_uint32), ("zeroed", C.c_int32)]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        L.orc_sad.restype = C.c_uint32
        L.orc_ssd16b.restype = C.c_uint32
        L.orc_tables_create.restype = C.c_void_p
        L.orc_tables_destroy.argtypes = [C.c_void_p]
        for f in ("orc_tables_scan",):
            getattr(L, f).restype = C.POINTER(C.c_uint32)
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("orc_tables_quant", "orc_tables_dequant"):
            getattr(L, f).restype = i32p
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_quant.argtypes = [C.c_void_p, i16p, i16p, i16p] + [C.c_int] * 8 + [C.POINTER(C.c_int)]
        L.orc_inv_quant.argtypes = [C.c_void_p, i16p, i16p] + [C.c_int] * 5
        L.orc_mv_cost.restype = C.c_uint32
        L.orc_mv_cost.argtypes = [C.POINTER(OrcMv), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_motion_estimation.argtypes = [C.POINTER(OrcMeIn), C.POINTER(OrcMeOut)]
        L.orc_mc_luma.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, OrcMv]
        L.orc_mc_chroma.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, OrcMv]
        L.orc_mc_luma_ex.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, OrcMv, C.c_int]
        L.orc_mc_chroma_ex.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, OrcMv, C.c_int]
        L.orc_encode_inter_tu.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, i16p, i16p, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.POINTER(OrcTuOut)]
        L.orc_encode_intra_tu.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, i16p, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_double, C.POINTER(OrcTuOut)]
        L.orc_adi_filter.argtypes = [i16p, i16p, C.c_int, C.c_int]
        L.orc_intra_predict.argtypes = [i16p, C.c_int, C.c_int, C.c_int, i16p, C.c_int]
        L.orc_intra_uses_filtered.argtypes = [C.c_int, C.c_int]
        L.orc_intra_mode_sads.argtypes = [i16p, C.c_int, i16p, C.c_int, C.POINTER(C.c_uint32)]
        L.orc_weighted_average.argtypes = [i16p, C.c_int, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int]
        L.orc_sao_ctu_stats.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_sao_offset_ctu.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.tables = L.orc_tables_create()
        _oracle = L
    return _oracle


_ref = None


def have_ref():
    build_oracle()
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "librefdrv.so"))


def ref():
    """(libhomer_ref, librefdrv) -- the compiled reference.  Only call when have_ref()."""
    global _ref
    if _ref is None:
        # librefdrv.so FIRST and global: oracle/ref_hooks.c interposes hmr_motion_estimation & co., which only works when the
        # harness precedes libhomer_ref.so (its dependency) in the lookup order of the calls made inside the reference
        ref_dir = os.environ.get("HB_REF_DIR", os.path.join(ORACLE_DIR, "_ref"))      # diagnostic builds of the reference (tools/ref_uninit_probe.sh)
        D = C.CDLL(os.path.join(ref_dir, "librefdrv.so"), mode=C.RTLD_GLOBAL)
        R = C.CDLL(os.path.join(ref_dir, "libhomer_ref.so"), mode=C.RTLD_GLOBAL)
        R.sse_aligned_sad.restype = C.c_uint32
        R.sse_aligned_ssd16b.restype = C.c_uint32
        R.sad.restype = C.c_uint32
        R.ssd16b.restype = C.c_uint32
        D.refdrv_open.restype = C.c_void_p
        D.refdrv_open.argtypes = [C.c_int] * 4
        D.refdrv_close.argtypes = [C.c_void_p]
        D.refdrv_sse_selected.argtypes = [C.c_void_p]
        D.refdrv_scan.restype = C.POINTER(C.c_uint32)
        D.refdrv_scan.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("refdrv_quant_table", "refdrv_dequant_table"):
            getattr(D, f).restype = i32p
            getattr(D, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        D.refdrv_quant.argtypes = [C.c_void_p, i16p, i16p, i16p] + [C.c_int] * 8 + [C.POINTER(C.c_int)]
        D.refdrv_quant_plainc.argtypes = [C.c_void_p, i16p, i16p] + [C.c_int] * 8 + [C.POINTER(C.c_int)]
        D.refdrv_inv_quant.argtypes = [C.c_void_p, i16p, i16p] + [C.c_int] * 5
        D.refdrv_motion_estimation.restype = C.c_uint32
        D.refdrv_motion_estimation.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int] + [C.c_int] * 5 + [
            C.c_int, i32p, C.c_int, i32p, C.c_int, C.c_double, C.c_uint, i32p]
        D.refdrv_mc_luma.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
        D.refdrv_mc_chroma.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
        D.refdrv_mc_luma_bi.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
        D.refdrv_mc_chroma_bi.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
        D.refdrv_encode_inter_tu.argtypes = [C.c_void_p, i16p, i16p] + [C.c_int] * 6 + [C.c_double, i16p, i16p,
                                                                                       C.POINTER(C.c_int)]
        D.refdrv_chroma_qp.argtypes = [C.c_void_p, C.c_int]
        D.refdrv_weighted_average.argtypes = [C.c_void_p, i16p, C.c_int, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int]
        D.refdrv_intra_predict.argtypes = [C.c_void_p, i16p, C.c_int, C.c_int, C.c_int, i16p]
        D.refdrv_adi_filter.argtypes = [C.c_void_p, i16p, i16p, C.c_int]
        D.refdrv_encode_lockstep.restype = C.c_long
        D.refdrv_encode_lockstep.argtypes = [C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             u8p, C.c_long, u8p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        _ref = (R, D)
    return _ref


_drv = {}


def refdrv(width=416, height=240, qp=32, sign_hiding=1):
    """a live reference encoder (its henc_thread_t is what quant / ME / the TQ chain need)"""
    key = (width, height, qp, sign_hiding)
    if key not in _drv:
        _, D = ref()
        h = D.refdrv_open(width, height, qp, sign_hiding)
        assert h, "reference SETCFG failed"
        _drv[key] = h
    return _drv[key]


def make_frame_pair(rng, w, h, pad, shift=(3, 2), noise=3.0):
    """Synthetic current / reference luma pair: block texture, blurred, panned, plus noise.
    Returns int16 planes of (h+2*pad) x (w+2*pad), reference border-replicated."""
    tex = rng.integers(0, 256, size=((h + 96) // 8 + 2, (w + 96) // 8 + 2)).astype(np.float64)
    tex = np.kron(tex, np.ones((8, 8)))
    k = np.ones(9) / 9.0
    tex = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, tex)
    tex = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, tex)

    def crop(ox, oy):
        f = tex[16 + oy:16 + oy + h, 16 + ox:16 + ox + w] + rng.normal(0, noise, size=(h, w))
        return np.clip(np.rint(f), 0, 255).astype(np.int16)

    cur = crop(shift[0], shift[1])
    ref_ = crop(0, 0)
    return np.pad(cur, pad, mode="edge"), np.pad(ref_, pad, mode="edge")


def chroma_qp(qp, offset=2):
    """chroma QP of a luma QP (chroma_scale_conversion_table, hmr_encoder_lib.c:2245)"""
    t = (C.c_uint8 * 58).in_dll(oracle(), "orc_chroma_qp_table")
    return int(t[min(max(qp + offset, 0), 57)])


# ---------------------------------------------------------------------------------------------------------------
# the frame-level pre-pass computed by the compiled reference's own functions (oracle/ref_driver.c: refdrv_prepass)
class _RpOut(C.Structure):
    _fields_ = [("me", C.c_void_p * 4), ("tu", (C.c_void_p * 3) * 5), ("coeff", (C.c_void_p * 3) * 5),
                ("recon", (C.c_void_p * 3) * 5), ("pred", (C.c_void_p * 3) * 4)]


ME_DT = np.dtype([("mvx", "<i4"), ("mvy", "<i4"), ("subx", "<i4"), ("suby", "<i4"), ("sad", "<u4"), ("n_probes", "<u4")])
TU_DT = np.dtype([("sum", "<i4"), ("ssd", "<u4"), ("ssd_zero", "<u4"), ("zeroed", "<i4")])
PASS_TU = (32, 32, 16, 8, 4)
_rp_handles = {}

SAO_DT = np.dtype([("eo_diff", "<i4", (4, 5)), ("eo_count", "<i4", (4, 5)), ("bo_diff", "<i4", (32,)), ("bo_count", "<i4", (32,))])


def oracle_sao_stats(rec, org, w, h):
    """rec / org: (y, u, v) uint8 planes.  Returns SAO_DT array [n_ctus, 3] from the C restatement."""
    O = oracle()
    cols, rows = (w + 63) // 64, (h + 63) // 64
    out = np.zeros((rows * cols, 3), SAO_DT)
    for c in range(3):
        pw, ph = (w, h) if c == 0 else (w // 2, h // 2)
        r16 = np.ascontiguousarray(np.pad(rec[c].astype(np.int16), 2)); o16 = np.ascontiguousarray(org[c].astype(np.int16))
        cs = 64 if c == 0 else 32
        for i in range(rows * cols):
            x0, y0 = (i % cols) * cs, (i // cols) * cs
            O.orc_sao_ctu_stats(ptr(r16.reshape(-1), 2 * (pw + 4) + 2), pw + 4, ptr(o16.reshape(-1)), pw, c, x0, y0, pw, ph, cs, out[i, c:c + 1].ctypes.data)
    return out


def random_sao_params(rng, w, h):
    """per CTU and component a type in -1..4 and the offsets the way sao_offset_t.offset holds them"""
    n = ((w + 63) // 64) * ((h + 63) // 64)
    types = rng.integers(-1, 5, (n, 3)).astype(np.int8)
    offs = np.zeros((n, 3, 32), np.int32)
    for i in range(n):
        for c in range(3):
            t = types[i, c]
            if 0 <= t < 4:
                offs[i, c, :5] = [rng.integers(0, 8), rng.integers(0, 8), 0, -rng.integers(0, 8), -rng.integers(0, 8)]
            elif t == 4:
                b0 = int(rng.integers(0, 29))
                offs[i, c, b0:b0 + 4] = rng.integers(-7, 8, 4)
    return types, offs


def oracle_sao_apply(src, w, h, types, offs):
    O = oracle()
    cols = (w + 63) // 64
    out = []
    for c in range(3):
        pw, ph = (w, h) if c == 0 else (w // 2, h // 2)
        cs = 64 if c == 0 else 32
        s16 = np.ascontiguousarray(np.pad(src[c].astype(np.int16), 2)); d16 = s16.copy()
        for i in range(len(types)):
            x0, y0 = (i % cols) * cs, (i // cols) * cs
            o = np.ascontiguousarray(offs[i, c], np.int32)
            O.orc_sao_offset_ctu(ptr(s16.reshape(-1), 2 * (pw + 4) + 2), pw + 4, ptr(d16.reshape(-1), 2 * (pw + 4) + 2), pw + 4, x0, y0, pw, ph, cs,
                                 int(types[i, c]), o.ctypes.data_as(C.POINTER(C.c_int)))
        out.append(d16[2:-2, 2:-2].astype(np.uint8))
    return out


def ref_sao_apply(src, w, h, types, offs):
    _, D = ref()
    hnd = refdrv()
    D.refdrv_sao_apply.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    src = [np.ascontiguousarray(p) for p in src]
    out = [np.zeros_like(p) for p in src]
    sp = (C.c_void_p * 3)(*[p.ctypes.data for p in src]); op = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    types = np.ascontiguousarray(types, np.int8); offs = np.ascontiguousarray(offs, np.int32)
    assert D.refdrv_sao_apply(hnd, sp, w, h, types.ctypes.data, offs.ctypes.data, op) == len(types)
    return out


def random_deblock_case(rng, w, h):
    """a random but consistent CU/TU quadtree per CTU with modes, cbf, QP and vectors per 4x4 unit (picture raster, 16 units per CTU
    and row), and a blocky picture so that the filters have something to do"""
    cols, rows = (w + 63) // 64, (h + 63) // 64
    uw, uh = cols * 16, rows * 16
    m = dict(cu=np.zeros((uh, uw), np.uint8), tu=np.zeros((uh, uw), np.uint8), intra=np.zeros((uh, uw), np.uint8), cbf=np.zeros((uh, uw), np.uint8),
             qp=np.full((uh, uw), 30, np.uint8), mv=np.zeros((uh, uw, 2), np.int16))

    def gen(x, y, size, depth):
        if x >= w or y >= h:
            return
        if size > 8 and (x + size > w or y + size > h or rng.random() < 0.55):
            for dy in (0, size // 2):
                for dx in (0, size // 2):
                    gen(x + dx, y + dy, size // 2, depth + 1)
            return
        u0, v0, n = x // 4, y // 4, size // 4
        m["cu"][v0:v0 + n, u0:u0 + n] = depth
        m["tu"][v0:v0 + n, u0:u0 + n] = 1 if size == 64 else int(rng.integers(0, 2))
        m["intra"][v0:v0 + n, u0:u0 + n] = rng.random() < 0.3
        m["cbf"][v0:v0 + n, u0:u0 + n] = rng.integers(0, 4, (n, n))
        m["qp"][v0:v0 + n, u0:u0 + n] = rng.integers(18, 46)
        m["mv"][v0:v0 + n, u0:u0 + n] = rng.integers(-6, 7, 2)
    for cy in range(rows):
        for cx in range(cols):
            gen(cx * 64, cy * 64, 64, 0)
    planes = []
    for (ww, hh, b) in ((w, h, 8), (w // 2, h // 2, 4), (w // 2, h // 2, 4)):
        p = np.clip(rng.normal(128, 30, (hh, ww)), 0, 255)
        mean = p.reshape(hh // b, b, ww // b, b).mean((1, 3))
        planes.append(np.clip(np.repeat(np.repeat(mean, b, 0), b, 1) + rng.integers(-3, 4, (hh, ww)), 0, 255).astype(np.uint8))
    return m, planes


_dbk_handles = {}


def ref_deblock(planes, w, h, m):
    """the reference's own deblocking (hmr_deblock_filter_cu per CTU and direction).  Returns (planes out, bs_ver, bs_hor, (cb, cr) qp offsets)"""
    _, D = ref()
    D.refdrv_open.restype = C.c_void_p
    D.refdrv_open.argtypes = [C.c_int] * 4
    if (w, h) not in _dbk_handles:
        _dbk_handles[(w, h)] = D.refdrv_open(w, h, 32, 1)
    hnd = _dbk_handles[(w, h)]
    D.refdrv_deblock.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.POINTER(C.c_void_p)]
    D.refdrv_pps_qp_offsets.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    planes = [np.ascontiguousarray(p) for p in planes]
    out = [np.zeros_like(p) for p in planes]
    bsv = np.zeros_like(m["cu"]); bsh = np.zeros_like(m["cu"])
    ip = (C.c_void_p * 3)(*[p.ctypes.data for p in planes]); op = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    a = {k: np.ascontiguousarray(v) for k, v in m.items()}
    n = D.refdrv_deblock(hnd, ip, w, h, a["cu"].ctypes.data, a["tu"].ctypes.data, a["intra"].ctypes.data, a["cbf"].ctypes.data, a["qp"].ctypes.data,
                         a["mv"].ctypes.data, bsv.ctypes.data, bsh.ctypes.data, op)
    assert n == ((w + 63) // 64) * ((h + 63) // 64)
    cb, cr = C.c_int(0), C.c_int(0)
    D.refdrv_pps_qp_offsets(hnd, C.byref(cb), C.byref(cr))
    return out, bsv, bsh, (cb.value, cr.value)


def oracle_deblock(planes, w, h, bsv, bsh, qp, offs, beta_off=0, tc_off=0):
    O = oracle()
    O.orc_deblock_picture.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5
    pl = [np.ascontiguousarray(p.astype(np.int16)) for p in planes]
    pp = (C.c_void_p * 3)(*[p.ctypes.data for p in pl]); st = (C.c_int * 3)(w, w // 2, w // 2)
    bsv = np.ascontiguousarray(bsv); bsh = np.ascontiguousarray(bsh); qp = np.ascontiguousarray(qp)
    O.orc_deblock_picture(pp, st, w, h, bsv.ctypes.data, bsh.ctypes.data, qp.ctypes.data, bsv.shape[1], offs[0], offs[1], beta_off, tc_off)
    return [p.astype(np.uint8) for p in pl]


def oracle_deblock_strengths(m, w, h):
    O = oracle()
    O.orc_deblock_strengths.argtypes = [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p] * 2
    a = {k: np.ascontiguousarray(v) for k, v in m.items()}
    ref_idx = np.ascontiguousarray(np.where(a["intra"] != 0, -1, 0).astype(np.int8))
    bsv = np.zeros_like(a["cu"]); bsh = np.zeros_like(a["cu"])
    O.orc_deblock_strengths(a["cu"].ctypes.data, a["tu"].ctypes.data, a["intra"].ctypes.data, a["cbf"].ctypes.data, ref_idx.ctypes.data,
                            a["mv"].ctypes.data, a["cu"].shape[1], w, h, bsv.ctypes.data, bsh.ctypes.data)
    return bsv, bsh


def ref_sao_stats(rec, org, w, h):
    """the same through the reference's table member get_sao_stats; returns SAO_DT array [n_ctus, 3]"""
    _, D = ref()
    hnd = refdrv()
    D.refdrv_sao_stats.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    rec = [np.ascontiguousarray(p) for p in rec]; org = [np.ascontiguousarray(p) for p in org]
    rp = (C.c_void_p * 3)(*[p.ctypes.data for p in rec]); op = (C.c_void_p * 3)(*[p.ctypes.data for p in org])
    n = ((w + 63) // 64) * ((h + 63) // 64)
    raw = np.zeros((n, 3, 5, 2, 32), np.int64)
    assert D.refdrv_sao_stats(hnd, rp, op, w, h, raw.ctypes.data) == n
    out = np.zeros((n, 3), SAO_DT)
    out["eo_diff"] = raw[:, :, 0:4, 0, 0:5]; out["eo_count"] = raw[:, :, 0:4, 1, 0:5]      # the reference offsets its pointers by 2: classes -2..2 -> 0..4
    out["bo_diff"] = raw[:, :, 4, 0, :]; out["bo_count"] = raw[:, :, 4, 1, :]
    assert (raw[:, :, 0:4, :, 5:] == 0).all()
    return out


def ref_sao_derive(rec, comp, sao_type, lam):
    """sao_derive_offsets + sao_invert_quant_offsets + sao_get_distortion of the reference on one SAO_DT record: (offsets[32], band, dist)"""
    _, D = ref()
    hnd = refdrv()
    D.refdrv_sao_derive.restype = C.c_int64
    D.refdrv_sao_derive.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int32)]
    diff = np.zeros(32, np.int64); count = np.zeros(32, np.int64)
    if sao_type < 4:
        diff[:5] = rec["eo_diff"][sao_type]; count[:5] = rec["eo_count"][sao_type]
    else:
        diff[:] = rec["bo_diff"]; count[:] = rec["bo_count"]
    off = np.zeros(32, np.int32); band = C.c_int32(0)
    dist = D.refdrv_sao_derive(hnd, diff.ctypes.data, count.ctypes.data, comp, sao_type, float(lam), off.ctypes.data, C.byref(band))
    return off, band.value, dist


def random_sao_stats(rng, n):
    """plausible statistics records: counts of a 64x64 CTU spread over the classes, differences of either sign, some empty classes"""
    st = np.zeros(n, SAO_DT)
    st["eo_count"] = rng.integers(0, 1500, (n, 4, 5)) * (rng.random((n, 4, 5)) > 0.15)
    st["eo_diff"] = (st["eo_count"] * rng.normal(0, 2.5, (n, 4, 5))).astype(np.int32)
    st["bo_count"] = rng.integers(0, 700, (n, 32)) * (rng.random((n, 32)) > 0.4)
    st["bo_diff"] = (st["bo_count"] * rng.normal(0, 3.0, (n, 32))).astype(np.int32)
    return st


def ref_intra_presearch(luma, jobs, adi, adi_off, n_threads=1):
    """35-mode SADs of every job through the reference's own functions; returns (seconds, sads (n,35) uint32)"""
    _, D = ref()
    D.refdrv_intra_presearch.restype = C.c_double
    D.refdrv_intra_presearch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    hs = _rp_handles.setdefault((32, 1), [])
    while len(hs) < n_threads:
        hh = D.refdrv_open(128, 128, 32, 1)
        assert hh
        hs.append(hh)
    handles = (C.c_void_p * n_threads)(*hs[:n_threads])
    luma = np.ascontiguousarray(luma, np.uint8); jobs = np.ascontiguousarray(jobs, np.int32)
    adi = np.ascontiguousarray(adi, np.int16); adi_off = np.ascontiguousarray(adi_off, np.int32)
    sads = np.zeros((len(jobs), 35), np.uint32)
    h, w = luma.shape
    secs = D.refdrv_intra_presearch(handles, n_threads, luma.ctypes.data, w, h, jobs.ctypes.data, len(jobs), adi.ctypes.data, adi_off.ctypes.data, sads.ctypes.data)
    return secs, sads


_rp_out_cache = {}


def ref_prepass(cur, ref_planes, w, h, qp=32, avg_dist=650.0, n_threads=1, band=(0, 0), sign_hiding=1, want_pred=True, reuse_outputs=False):
    """cur / ref_planes: (y, u, v) uint8 planes.  Returns (seconds, dict) with the GPU library's output layouts."""
    _, D = ref()
    D.refdrv_prepass.restype = C.c_double
    D.refdrv_prepass.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                 C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(_RpOut)]
    D.refdrv_prepass_num_tus.argtypes = [C.c_int] * 6
    key = (qp, sign_hiding)
    hs = _rp_handles.setdefault(key, [])
    while len(hs) < n_threads:
        hh = D.refdrv_open(128, 128, qp, sign_hiding)      # per-thread scratch is CTU sized; the picture size is irrelevant
        assert hh
        hs.append(hh)
    handles = (C.c_void_p * n_threads)(*hs[:n_threads])
    cur = [np.ascontiguousarray(p) for p in cur]
    ref_planes = [np.ascontiguousarray(p) for p in ref_planes]
    cp = (C.c_void_p * 3)(*[p.ctypes.data for p in cur])
    rp = (C.c_void_p * 3)(*[p.ctypes.data for p in ref_planes])
    # output arrays are allocated once per shape and reused (bench.py times this call in a loop: ~80 MB of np.zeros per frame
    # would be timed as the reference's work otherwise); results of a previous call with the same shape are overwritten
    ck = (w, h, tuple(band), bool(want_pred))
    if reuse_outputs and ck in _rp_out_cache:
        out, res = _rp_out_cache[ck]
        for a in res["me"]:
            a["sad"] = 0xFFFFFFFF
        secs = D.refdrv_prepass(handles, n_threads, cp, rp, w, h, qp, avg_dist, band[0], band[1], C.byref(out))
        return secs, res
    out = _RpOut()
    res = {"me": [], "tu": {}, "coeff": {}, "recon": [], "pred": []}
    cc, cr = (w + 63) // 64, (h + 63) // 64
    for d in range(4):
        s = 64 >> d
        a = np.zeros(cc * (64 // s) * cr * (64 // s), ME_DT)
        a["sad"] = 0xFFFFFFFF
        res["me"].append(a)
        out.me[d] = a.ctypes.data
        if want_pred:
            pl = [np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8)]
            res["pred"].append(pl)
            for c in range(3):
                out.pred[d][c] = pl[c].ctypes.data
    for p in range(5):
        pl = [np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8)]
        res["recon"].append(pl)
        for c in range(3):
            out.recon[p][c] = pl[c].ctypes.data
            n = D.refdrv_prepass_num_tus(w, h, band[0], band[1], p, c)
            t = PASS_TU[p] // (2 if c else 1)
            res["tu"][(p, c)] = np.zeros(n, TU_DT)
            res["coeff"][(p, c)] = np.zeros((n, t, t), np.int16)
            if n:
                out.tu[p][c] = res["tu"][(p, c)].ctypes.data
                out.coeff[p][c] = res["coeff"][(p, c)].ctypes.data
    if reuse_outputs:
        _rp_out_cache[ck] = (out, res)
    secs = D.refdrv_prepass(handles, n_threads, cp, rp, w, h, qp, avg_dist, band[0], band[1], C.byref(out))
    return secs, res


def amvp_jobs(w, h):
    """every 2Nx2N PU of every size inside the picture: (x, y, size)"""
    return np.array([(x, y, s) for s in (64, 32, 16, 8) for y in range(0, h - s + 1, s) for x in range(0, w - s + 1, s)], np.int32)


def ref_amvp(w, h, m, jobs):
    """the reference's own get_amvp_candidates on CTU descriptions built from the unit maps: (n, 4) int32"""
    _, D = ref()
    D.refdrv_open.restype = C.c_void_p
    D.refdrv_open.argtypes = [C.c_int] * 4
    if (w, h) not in _dbk_handles:
        _dbk_handles[(w, h)] = D.refdrv_open(w, h, 32, 1)
    D.refdrv_amvp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    inter = np.ascontiguousarray(1 - m["intra"].astype(np.uint8)); mv = np.ascontiguousarray(m["mv"]); jobs = np.ascontiguousarray(jobs, np.int32)
    out = np.zeros((len(jobs), 4), np.int32)
    n = D.refdrv_amvp(_dbk_handles[(w, h)], w, h, inter.ctypes.data, mv.ctypes.data, jobs.ctypes.data, len(jobs), out.ctypes.data)
    assert n == len(jobs), n
    return out


def oracle_amvp(w, h, m, jobs):
    O = oracle()
    O.orc_amvp_candidates.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]
    inter = np.ascontiguousarray(1 - m["intra"].astype(np.uint8)); mv = np.ascontiguousarray(m["mv"])
    out = np.zeros((len(jobs), 4), np.int32)
    for i, (x, y, s) in enumerate(jobs):
        O.orc_amvp_candidates(inter.ctypes.data, mv.ctypes.data, inter.shape[1], w, h, int(x), int(y), int(s), out[i].ctypes.data)
    return out


def ref_merge(w, h, m, jobs, max_cands):
    """the reference's own get_merge_mvp_candidates: (n, max_cands, 2) int32"""
    _, D = ref()
    D.refdrv_open.restype = C.c_void_p
    D.refdrv_open.argtypes = [C.c_int] * 4
    if (w, h) not in _dbk_handles:
        _dbk_handles[(w, h)] = D.refdrv_open(w, h, 32, 1)
    D.refdrv_amvp_or_merge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    inter = np.ascontiguousarray(1 - m["intra"].astype(np.uint8)); mv = np.ascontiguousarray(m["mv"]); jobs = np.ascontiguousarray(jobs, np.int32)
    out = np.zeros((len(jobs), max_cands, 2), np.int32)
    n = D.refdrv_amvp_or_merge(_dbk_handles[(w, h)], w, h, inter.ctypes.data, mv.ctypes.data, jobs.ctypes.data, len(jobs), max_cands, out.ctypes.data)
    assert n == len(jobs), n
    return out


def oracle_merge(w, h, m, jobs, max_cands):
    O = oracle()
    O.orc_merge_candidates.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    inter = np.ascontiguousarray(1 - m["intra"].astype(np.uint8)); mv = np.ascontiguousarray(m["mv"])
    out = np.zeros((len(jobs), max_cands, 2), np.int32)
    for i, (x, y, s) in enumerate(jobs):
        O.orc_merge_candidates(inter.ctypes.data, mv.ctypes.data, inter.shape[1], w, h, int(x), int(y), int(s), max_cands, out[i].ctypes.data)
    return out


def random_b_motion(rng, m, n_l0=2, n_l1=2):
    """two-list motion for the units of a random_deblock_case: per CU a prediction direction (list 0, list 1 or both), reference indices
    and vectors, few distinct values so that equal / swapped / nearly equal motion on the two sides of an edge all occur; the pictures
    the indices name overlap between the lists"""
    uh, uw = m["cu"].shape
    ref0 = np.full((uh, uw), -1, np.int8); ref1 = np.full((uh, uw), -1, np.int8)
    mv0 = np.zeros((uh, uw, 2), np.int16); mv1 = np.zeros((uh, uw, 2), np.int16)
    for uy in range(0, uh, 2):
        for ux in range(0, uw, 2):
            size = max(64 >> int(m["cu"][uy, ux]), 8) // 4
            if (uy % size) or (ux % size):
                continue
            d = int(rng.integers(1, 4))
            if d & 1:
                ref0[uy:uy + size, ux:ux + size] = rng.integers(0, n_l0); mv0[uy:uy + size, ux:ux + size] = rng.integers(-1, 2, 2) * 5
            if d & 2:
                ref1[uy:uy + size, ux:ux + size] = rng.integers(0, n_l1); mv1[uy:uy + size, ux:ux + size] = rng.integers(-1, 2, 2) * 5
    ref0[m["intra"] != 0] = -1; ref1[m["intra"] != 0] = -1
    pic_l0 = np.array([3, 5][:n_l0], np.int32); pic_l1 = np.array([5, 3][:n_l1], np.int32)       # the lists name the same two pictures, in opposite order
    return ref0, mv0, ref1, mv1, pic_l0, pic_l1


def ref_deblock_strengths_b(w, h, m, ref0, mv0, ref1, mv1, pic_l0, pic_l1):
    _, D = ref()
    D.refdrv_open.restype = C.c_void_p
    D.refdrv_open.argtypes = [C.c_int] * 4
    if (w, h) not in _dbk_handles:
        _dbk_handles[(w, h)] = D.refdrv_open(w, h, 32, 1)
    D.refdrv_deblock_strengths_b.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    a = {k: np.ascontiguousarray(v) for k, v in m.items()}
    arrs = [np.ascontiguousarray(x) for x in (ref0, mv0, ref1, mv1, pic_l0, pic_l1)]
    bsv = np.zeros_like(a["cu"]); bsh = np.zeros_like(a["cu"])
    n = D.refdrv_deblock_strengths_b(_dbk_handles[(w, h)], w, h, a["cu"].ctypes.data, a["tu"].ctypes.data, a["intra"].ctypes.data, a["cbf"].ctypes.data,
                                     arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, arrs[4].ctypes.data, len(pic_l0),
                                     arrs[5].ctypes.data, len(pic_l1), bsv.ctypes.data, bsh.ctypes.data)
    assert n > 0, n
    return bsv, bsh


def oracle_deblock_strengths_b(w, h, m, ref0, mv0, ref1, mv1, pic_l0, pic_l1):
    O = oracle()
    O.orc_deblock_strengths_b.argtypes = [C.c_void_p] * 10 + [C.c_int] * 3 + [C.c_void_p] * 2
    a = {k: np.ascontiguousarray(v) for k, v in m.items()}
    arrs = [np.ascontiguousarray(x) for x in (ref0, mv0, ref1, mv1, pic_l0, pic_l1)]
    bsv = np.zeros_like(a["cu"]); bsh = np.zeros_like(a["cu"])
    O.orc_deblock_strengths_b(a["cu"].ctypes.data, a["tu"].ctypes.data, a["intra"].ctypes.data, a["cbf"].ctypes.data, arrs[0].ctypes.data, arrs[1].ctypes.data,
                              arrs[2].ctypes.data, arrs[3].ctypes.data, arrs[4].ctypes.data, arrs[5].ctypes.data, bsv.shape[1], w, h, bsv.ctypes.data, bsh.ctypes.data)
    return bsv, bsh
