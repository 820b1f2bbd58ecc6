"""Whole-encode helpers shared by the CPU and GPU tests: the reference's encoder in lock step (oracle/ref_driver.c) with an optional
hook that installs a replacement function table and / or the CU-granularity hooks of oracle/ref_hooks.c."""
import ctypes as C

import numpy as np

from homerhevc_b200 import synth
from _oracle import ref

HOOK_COUNTERS = "frames p_frames me me_fwd mc mc_cached mc_fwd tq tq_cached tq_fwd tq_stale errors".split()


class TableHook(C.Structure):            # user data of refdrv_install_gpu_table / refdrv_install_shadow_table
    _fields_ = [("lib", C.c_void_p), ("which", C.c_int)]


class CuHookCfg(C.Structure):            # user data of refdrv_install_cu_hooks (cu_hook_cfg, oracle/ref_hooks.c)
    _fields_ = [("lib", C.c_void_p), ("which", C.c_int), ("batch_tq", C.c_int)]


class ShadowReport(C.Structure):         # shadow_report, oracle/ref_shadow.c
    _fields_ = [("calls", C.c_long), ("mismatches", C.c_long), ("first_fn", C.c_int), ("args", C.c_int * 12), ("first_call", C.c_long),
                ("first_at", C.c_int), ("cpu_val", C.c_int32), ("gpu_val", C.c_int32), ("per_fn", C.c_long * 12)]


def make_yuv(w, h, nf, seed=21):
    clip = synth.make_clip(w, h, nf, seed=seed)
    return np.concatenate([np.concatenate([p.reshape(-1) for p in f]) for f in clip])


def encode(w, h, yuv, nf, hook=None, user=None, force_intra=0, perf=-1, qp=32, sign_hiding=1):
    """-> (bitstream bytes, reconstruction, seconds).  hook: address of a refdrv_table_hook, user: ctypes object passed to it."""
    _, D = ref()
    bs = np.zeros(32 << 20, np.uint8); rec = np.zeros(yuv.size, np.uint8); secs = C.c_double(0)
    n = D.refdrv_encode_lockstep(w, h, nf, yuv.ctypes.data_as(C.POINTER(C.c_uint8)), qp, sign_hiding, force_intra, perf,
                                 bs.ctypes.data_as(C.POINTER(C.c_uint8)), bs.size, rec.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 hook, C.cast(C.pointer(user), C.c_void_p) if user is not None else None, C.byref(secs))
    assert n > 0, "reference encode failed"
    return bytes(bs[:n]), rec, secs.value


def hook_addr(name):
    _, D = ref()
    f = getattr(D, name + "_addr")
    f.restype = C.c_void_p
    return C.c_void_p(f())


def cu_hooks_off():
    """switch the CU hooks off and return their counters as a dict"""
    _, D = ref()
    cnt = (C.c_long * len(HOOK_COUNTERS))()
    D.refdrv_cu_hooks_off(cnt)
    return dict(zip(HOOK_COUNTERS, [int(v) for v in cnt]))


def shadow_report():
    _, D = ref()
    r = ShadowReport()
    D.refdrv_shadow_report(C.byref(r))
    return r


def describe_mismatch(w, h, a_bs, a_rec, b_bs, b_rec):
    first = next((i for i, (x, y) in enumerate(zip(a_bs, b_bs)) if x != y), min(len(a_bs), len(b_bs)))
    bad = np.flatnonzero(a_rec != b_rec)
    frame = int(bad[0]) // (w * h * 3 // 2) if bad.size else -1
    return f"bitstreams differ from byte {first} ({len(a_bs)} vs {len(b_bs)} bytes), {bad.size} reconstructed samples differ, first in frame {frame}"


class Capture(C.Structure):              # refdrv_capture, oracle/ref_hooks.c
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("frame", C.c_int32), ("n_seen", C.c_int32), ("recon", C.c_void_p * 3),
                ("pred_depth", C.c_void_p), ("part_size", C.c_void_p), ("mode_y", C.c_void_p), ("mode_c", C.c_void_p), ("tr_idx", C.c_void_p),
                ("qp", C.c_void_p), ("pred_mode", C.c_void_p), ("cbf", C.c_void_p * 3), ("coeff", C.c_void_p), ("slice_type", C.c_int32), ("slice_qp", C.c_int32)]


def encode_and_capture(w, h, yuv, nf, frame=0, force_intra=0, qp=32, sign_hiding=1):
    """the reference's own encode of `nf` frames; of picture `frame`: its decisions per 4x4 unit, its levels and its reconstruction BEFORE
    the in-loop filters, taken CTU by CTU as the encoder finishes them.  -> dict of numpy arrays"""
    _, D = ref()
    cols, rows = (w + 63) // 64, (h + 63) // 64
    a = {k: np.zeros((rows * 16, cols * 16), np.uint8) for k in ("pred_depth", "part_size", "mode_y", "mode_c", "tr_idx", "qp", "pred_mode", "cbf_y", "cbf_u", "cbf_v")}
    rec = [np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8)]
    coeff = np.zeros((cols * rows, 64 * 64 + 2 * 32 * 32), np.int16)
    c = Capture()
    c.width, c.height, c.frame = w, h, frame
    for i in range(3):
        c.recon[i] = rec[i].ctypes.data; c.cbf[i] = a["cbf_" + "yuv"[i]].ctypes.data
    for k in ("pred_depth", "part_size", "mode_y", "mode_c", "tr_idx", "qp", "pred_mode"):
        setattr(c, k, a[k].ctypes.data)
    c.coeff = coeff.ctypes.data
    D.refdrv_capture_set.argtypes = [C.c_void_p]
    D.refdrv_capture_set(C.byref(c))
    try:
        bs, out, _ = encode(w, h, yuv, nf, force_intra=force_intra, qp=qp, sign_hiding=sign_hiding)
    finally:
        D.refdrv_capture_set(None)
    assert c.n_seen == cols * rows, (c.n_seen, cols * rows)
    a.update(recon=rec, coeff=coeff, slice_type=int(c.slice_type), slice_qp=int(c.slice_qp), filtered=out)
    return a
