"""ctypes bindings of libhomer_b200.so (see include/homer_b200.h).  Host-side mirror only -- no computation here."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ME_PEL, ME_HALF, ME_QUARTER = 1, 2, 4
REG_DCT = 65535

i16p = C.POINTER(C.c_int16)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)


class HbError(RuntimeError):
    pass


class Mv(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class MeJob(C.Structure):
    """one hmr_motion_estimation call (hmr_motion_inter.c:1404)"""
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("qp", C.c_int32),
                ("n_amvp", C.c_int32), ("amvp", Mv * 2), ("n_start", C.c_int32), ("start", Mv * 3),
                ("parent", C.c_int32)]


class MeResult(C.Structure):
    _fields_ = [("mv", Mv), ("subpix", Mv), ("sad", C.c_uint32), ("n_probes", C.c_uint32)]


class McJob(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("mv", Mv)]


class McBiJob(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("mv0", Mv), ("mv1", Mv)]


class TuJob(C.Structure):
    _fields_ = [("comp", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("qp", C.c_int32)]


class IntraTuJob(C.Structure):
    _fields_ = [("comp", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("qp", C.c_int32), ("scan_mode", C.c_int32)]


class IntraJob(C.Structure):
    _fields_ = [("comp", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("size", C.c_int32), ("mode", C.c_int32), ("filtered", C.c_int32)]


class TuResult(C.Structure):
    _fields_ = [("sum", C.c_int32), ("ssd", C.c_uint32), ("ssd_zero", C.c_uint32), ("zeroed", C.c_int32)]


class TqParams(C.Structure):
    _fields_ = [("is_islice", C.c_int32), ("sign_hiding", C.c_int32), ("avg_dist", C.c_double),
                ("chroma_weight", C.c_double)]


class QuantEnv(C.Structure):
    _fields_ = [("is_islice", C.c_int32), ("sign_hiding", C.c_int32), ("max_cu_size_shift", C.c_int32),
                ("bit_depth", C.c_int32), ("delta_u", i16p)]


class PrepassCfg(C.Structure):
    _fields_ = [("qp", C.c_int32), ("chroma_qp_offset", C.c_int32), ("sign_hiding", C.c_int32),
                ("is_islice", C.c_int32), ("me_action", C.c_int32), ("use_graph", C.c_int32),
                ("band_ctu_row0", C.c_int32), ("band_ctu_rows", C.c_int32), ("compact_tables", C.c_int32), ("subpel_per_pu", C.c_int32), ("me_staged_window", C.c_int32), ("me_per_depth", C.c_int32)]


# numpy views of the compact wire records (hb_me_result_c / hb_tu_result_c, 12 bytes each)
ME_COMPACT_DT = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("sad", "<u4"), ("n_probes", "<u2"), ("subx", "i1"), ("suby", "i1")])
TU_COMPACT_DT = np.dtype([("ssd", "<u4"), ("ssd_zero", "<u4"), ("sum_zeroed", "<u4")])
CU_COST_DT = np.dtype([("ssd", "<u4"), ("sum", "<u4"), ("cbf", "<u2"), ("reserved", "<u2")])      # hb_cu_cost, compact_tables = 2


class LowLevelFuncs(C.Structure):
    """member-for-member image of low_level_funcs_t (19 pointers, hmr_private.h:1063-1092)"""
    _fields_ = [(n, C.c_void_p) for n in (
        "sse_copy_16_16", "sse_copy_16_8", "sse_copy_8_16", "sad", "ssd16b", "predict", "reconst",
        "modified_variance", "create_intra_planar_prediction", "create_intra_angular_prediction",
        "interpolate_luma_m_compensation", "interpolate_chroma_m_compensation", "interpolate_luma_m_estimation",
        "weighted_average_motion", "quant", "inv_quant", "transform", "itransform", "get_sao_stats")]


def library_path():
    return os.path.join(_HERE, "libhomer_b200.so")


def build_library(verbose=False):
    """compile every CUDA/C source for sm_100a into libhomer_b200.so (nvcc cross-compiles without a GPU)"""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise HbError("building libhomer_b200.so failed")
    return library_path()


_lib = None


def load_library():
    """dlopen the CUDA library.  Fails loudly when it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise HbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(the hot path is CUDA only, there is no CPU fallback)")
    L = C.CDLL(path)
    L.hb_last_error.restype = C.c_char_p
    L.hb_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.hb_ctx_destroy.argtypes = [C.c_void_p]
    L.hb_ctx_sync.argtypes = [C.c_void_p]
    L.hb_ctx_wait.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_ctx_stream.restype = C.c_void_p
    L.hb_ctx_stream.argtypes = [C.c_void_p]
    L.hb_ctx_launch_count.restype = C.c_uint64
    L.hb_ctx_launch_count.argtypes = [C.c_void_p]
    L.hb_timer_begin.argtypes = [C.c_void_p]
    L.hb_timer_end.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.hb_pinned_alloc.restype = C.c_void_p
    L.hb_pinned_alloc.argtypes = [C.c_size_t]
    L.hb_pinned_free.argtypes = [C.c_void_p]
    # section A
    L.hb_sad.restype = C.c_uint32
    L.hb_sad.argtypes = [i16p, C.c_uint32, i16p, C.c_uint32, C.c_int]
    L.hb_ssd16b.restype = C.c_uint32
    L.hb_ssd16b.argtypes = [i16p, C.c_uint32, i16p, C.c_uint32, C.c_int]
    L.hb_predict.restype = None
    L.hb_predict.argtypes = [i16p, C.c_int, i16p, C.c_int, i16p, C.c_int, C.c_int]
    L.hb_reconst.restype = None
    L.hb_reconst.argtypes = [i16p, C.c_int, i16p, C.c_int, i16p, C.c_int, C.c_int]
    L.hb_weighted_average_motion.restype = None
    L.hb_weighted_average_motion.argtypes = [i16p, C.c_int, i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
    for f in (L.hb_interpolate_luma, L.hb_interpolate_chroma):
        f.restype = None
        f.argtypes = [i16p, C.c_int, i16p, C.c_int] + [C.c_int] * 6
    L.hb_transform.restype = None
    L.hb_transform.argtypes = [C.c_int, i16p, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint16, i16p]
    L.hb_itransform.restype = None
    L.hb_itransform.argtypes = [C.c_int, i16p, i16p, C.c_int, C.c_int, C.c_int, C.c_uint, i16p]
    L.hb_quant.restype = None
    L.hb_quant.argtypes = [C.POINTER(QuantEnv), i16p, i16p] + [C.c_int] * 5 + [C.POINTER(C.c_int)] + [C.c_int] * 3
    L.hb_inv_quant.restype = None
    L.hb_inv_quant.argtypes = [C.POINTER(QuantEnv), i16p, i16p] + [C.c_int] * 6
    L.hb_default_ctx.restype = C.c_void_p
    L.hb_fill_low_level_funcs.restype = None
    L.hb_fill_low_level_funcs.argtypes = [C.POINTER(LowLevelFuncs)]
    # section B
    L.hb_frame_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.hb_frame_destroy.argtypes = [C.c_void_p]
    L.hb_frame_upload_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.hb_frame_upload_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.hb_frame_download_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.hb_frame_export_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.hb_frame_import_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.hb_frame_pad.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_frame_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_amvp_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.hb_merge_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.hb_frame_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.hb_frame_pull_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.hb_ipc_event_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    L.hb_ipc_event_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.hb_ipc_event_record.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_ipc_event_wait.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_ipc_event_destroy.argtypes = [C.c_void_p]
    L.hb_ipc_event_destroy.restype = None
    # section C
    L.hb_me_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MeJob), C.c_int, C.POINTER(MeResult), C.c_int,
                               C.c_double, C.c_int, C.POINTER(MeResult)]
    L.hb_mc_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(McJob), C.c_int]
    L.hb_mc_predict_bi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(McBiJob), C.c_int]
    L.hb_deblock_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.hb_deblock_frame_units.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_sao_stats_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_sao_apply_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_sao_derive_offsets.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.hb_sao_decide_standin.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    L.hb_merge_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(McJob), C.c_int, C.c_int, C.c_int, C.POINTER(TqParams), C.c_void_p]
    L.hb_tq_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TuJob), C.c_int,
                               C.POINTER(TqParams), i16p, C.POINTER(TuResult)]
    L.hb_tq_encode_intra.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(IntraTuJob), C.c_int, C.c_int, C.c_int, C.c_double,
                                     i16p, C.POINTER(TuResult)]
    L.hb_intra_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(IntraJob), C.c_int, i16p, C.POINTER(C.c_uint32)]
    L.hb_create_intra_planar_prediction.restype = None
    L.hb_create_intra_planar_prediction.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int]
    L.hb_create_intra_angular_prediction.restype = None
    L.hb_create_intra_angular_prediction.argtypes = [i16p, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int]
    # section D
    L.hb_prepass_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PrepassCfg), C.POINTER(C.c_void_p)]
    L.hb_prepass_destroy.argtypes = [C.c_void_p]
    L.hb_prepass_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.hb_prepass_num_pus.argtypes = [C.c_void_p, C.c_int]
    L.hb_prepass_run_profiled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_float), C.c_int]
    L.hb_prepass_kernel_name.restype = C.c_char_p
    L.hb_prepass_kernel_name.argtypes = [C.c_void_p, C.c_int]
    L.hb_prepass_num_tus.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hb_prepass_tu_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hb_prepass_tu_xy.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p]
    L.hb_prepass_fetch_me.argtypes = [C.c_void_p, C.c_int, C.POINTER(MeResult)]
    L.hb_prepass_fetch_tu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(TuResult)]
    L.hb_prepass_fetch_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, i16p]
    L.hb_prepass_fetch_recon.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.hb_prepass_fetch_all.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hb_prepass_output_bytes.restype = C.c_size_t
    L.hb_prepass_output_bytes.argtypes = [C.c_void_p]
    L.hb_prepass_tables_bytes.restype = C.c_size_t
    L.hb_prepass_tables_bytes.argtypes = [C.c_void_p]
    L.hb_prepass_fetch_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hb_prepass_num_ctus.argtypes = [C.c_void_p]
    L.hb_prepass_select.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.hb_prepass_gather_bytes.restype = C.c_size_t
    L.hb_prepass_gather_bytes.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_prepass_finalise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hb_prepass_fetch_units.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_prepass_frame_begin_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_double, C.c_void_p, C.c_size_t]
    L.hb_prepass_frame_finish_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(C.c_double), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
    L.hb_sao_candidates_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    L.hb_sao_decide_from_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    L.hb_prepass_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hb_prepass_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hb_prepass_frame_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double, C.c_void_p, C.c_size_t]
    L.hb_prepass_frame_finish.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hb_prepass_pred.restype = C.c_void_p
    L.hb_prepass_pred.argtypes = [C.c_void_p, C.c_int]
    L.hb_prepass_recon.restype = C.c_void_p
    L.hb_prepass_recon.argtypes = [C.c_void_p, C.c_int]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise HbError(f"{what}: rc={rc}: {load_library().hb_last_error().decode(errors='replace')}")


def _p16(a, off=0):
    assert a.dtype == np.int16 and a.flags["C_CONTIGUOUS"]
    return C.cast(a.ctypes.data + 2 * off, i16p)


class _LowLevel:
    """the per-call drop-ins under the reference's own member names (low_level_funcs_t); numpy int16 in/out,
    `off` arguments are element offsets into flat arrays so callers can address the interior of padded windows"""

    def sad(self, src, src_stride, pred, pred_stride, size, src_off=0, pred_off=0):
        return load_library().hb_sad(_p16(src, src_off), src_stride, _p16(pred, pred_off), pred_stride, size)

    def ssd16b(self, src, src_stride, pred, pred_stride, size, src_off=0, pred_off=0):
        return load_library().hb_ssd16b(_p16(src, src_off), src_stride, _p16(pred, pred_off), pred_stride, size)

    def predict(self, orig, orig_stride, pred, pred_stride, residual, residual_stride, size):
        load_library().hb_predict(_p16(orig), orig_stride, _p16(pred), pred_stride, _p16(residual), residual_stride, size)

    def reconst(self, pred, pred_stride, residual, residual_stride, decoded, decoded_stride, size):
        load_library().hb_reconst(_p16(pred), pred_stride, _p16(residual), residual_stride, _p16(decoded), decoded_stride, size)

    def interpolate_luma(self, ref, ref_stride, dst, dst_stride, fraction, width, height, is_vertical, is_first, is_last,
                         ref_off=0):
        load_library().hb_interpolate_luma(_p16(ref, ref_off), ref_stride, _p16(dst), dst_stride, fraction, width, height,
                                           is_vertical, is_first, is_last)

    def interpolate_chroma(self, ref, ref_stride, dst, dst_stride, fraction, width, height, is_vertical, is_first, is_last,
                           ref_off=0):
        load_library().hb_interpolate_chroma(_p16(ref, ref_off), ref_stride, _p16(dst), dst_stride, fraction, width, height,
                                             is_vertical, is_first, is_last)

    def transform(self, bit_depth, block, coeff, block_stride, n, mode=REG_DCT):
        lg = int(n).bit_length() - 1
        load_library().hb_transform(bit_depth, _p16(block), _p16(coeff), block_stride, n, n, lg, lg, mode, None)

    def itransform(self, bit_depth, block, coeff, block_stride, n, mode=REG_DCT):
        load_library().hb_itransform(bit_depth, _p16(block), _p16(coeff), block_stride, n, n, mode, None)

    def quant(self, env, src, dst, scan_mode, depth, comp, cu_mode, is_intra, cu_size, per, rem):
        s = C.c_int(0)
        load_library().hb_quant(C.byref(env), _p16(src), _p16(dst), scan_mode, depth, comp, cu_mode, is_intra, C.byref(s),
                                cu_size, per, rem)
        return s.value

    def inv_quant(self, env, src, dst, depth, comp, is_intra, cu_size, per, rem):
        load_library().hb_inv_quant(C.byref(env), _p16(src), _p16(dst), depth, comp, is_intra, cu_size, per, rem)


lowlevel = _LowLevel()


class Context:
    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        _check(self.L.hb_ctx_create(C.byref(h), device), "hb_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.hb_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        _check(self.L.hb_ctx_sync(self.h), "hb_ctx_sync")

    def wait(self, other):
        """work queued on this context from now on starts after everything queued so far on `other`"""
        _check(self.L.hb_ctx_wait(self.h, other.h), "hb_ctx_wait")

    def launch_count(self):
        return int(self.L.hb_ctx_launch_count(self.h))

    def stream_ptr(self):
        """the cudaStream_t the context queues its work on (for torch.cuda.ExternalStream)"""
        return int(self.L.hb_ctx_stream(self.h) or 0)

    def timer_begin(self):
        _check(self.L.hb_timer_begin(self.h), "hb_timer_begin")

    def timer_end(self):
        ms = C.c_float(0)
        _check(self.L.hb_timer_end(self.h, C.byref(ms)), "hb_timer_end")
        return ms.value

    def pinned(self, nbytes):
        """numpy uint8 view of pinned host memory (freed with the process)"""
        p = self.L.hb_pinned_alloc(nbytes)
        if not p:
            raise HbError("hb_pinned_alloc failed: " + self.L.hb_last_error().decode())
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(nbytes,))

    # ---- section C
    def me_search(self, cur, ref, jobs, avg_dist, action=7, parent_results=None):
        n = len(jobs)
        arr = (MeJob * n)(*jobs)
        out = (MeResult * n)()
        npar = len(parent_results) if parent_results is not None else 0
        par = (MeResult * npar)(*parent_results) if npar else None
        _check(self.L.hb_me_search(self.h, cur.h, ref.h, arr, n, par, npar, avg_dist, action, out), "hb_me_search")
        return list(out)

    def me_search_field(self, cur, ref, units, jobs, avg_dist, action=7):
        """hb_me_search with the AMVP predictors of every job taken on the device from the per-unit motion field"""
        units = np.ascontiguousarray(units, UNIT_INFO_DT)
        n = len(jobs)
        arr = (MeJob * n)(*jobs)
        out = (MeResult * n)()
        self.L.hb_me_search_field.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(MeJob), C.c_int, C.POINTER(MeResult), C.c_int,
                                              C.c_double, C.c_int, C.POINTER(MeResult)]
        _check(self.L.hb_me_search_field(self.h, cur.h, ref.h, units.ctypes.data, units.shape[1], arr, n, None, 0, avg_dist, action, out), "hb_me_search_field")
        return list(out)

    def mc_predict(self, ref, pred, jobs):
        n = len(jobs)
        arr = (McJob * n)(*jobs)
        _check(self.L.hb_mc_predict(self.h, ref.h, pred.h, arr, n), "hb_mc_predict")

    def mc_predict_bi(self, ref0, ref1, pred, jobs):
        n = len(jobs)
        arr = (McBiJob * n)(*jobs)
        _check(self.L.hb_mc_predict_bi(self.h, ref0.h, ref1.h, pred.h, arr, n), "hb_mc_predict_bi")

    def merge_eval(self, cur, ref, pred, recon, cands, qp, chroma_qp_offset, params):
        """merge / skip candidates (non-overlapping blocks): returns an (n,) array {dist_coded, sum, dist_skip, cbf}"""
        n = len(cands)
        arr = (McJob * n)(*cands)
        out = np.zeros(n, np.dtype([("dist_coded", "<u4"), ("sum", "<i4"), ("dist_skip", "<u4"), ("cbf", "<i4")]))
        _check(self.L.hb_merge_eval(self.h, cur.h, ref.h, pred.h, recon.h, arr, n, qp, chroma_qp_offset, C.byref(params), out.ctypes.data), "hb_merge_eval")
        return out

    def tq_encode(self, cur, pred, recon, jobs, params):
        n = len(jobs)
        arr = (TuJob * n)(*jobs)
        total = sum(j.size * j.size for j in jobs)
        coeffs = np.zeros(total, dtype=np.int16)
        res = (TuResult * n)()
        _check(self.L.hb_tq_encode(self.h, cur.h, pred.h, recon.h, arr, n, C.byref(params), _p16(coeffs), res), "hb_tq_encode")
        return coeffs, list(res)


def _tq_encode_intra(self, cur, pred, recon, jobs, is_islice, sign_hiding, chroma_weight):
    n = len(jobs)
    arr = (IntraTuJob * n)(*jobs)
    coeffs = np.zeros(sum(j.size * j.size for j in jobs), dtype=np.int16)
    res = (TuResult * n)()
    _check(self.L.hb_tq_encode_intra(self.h, cur.h, pred.h, recon.h, arr, n, is_islice, sign_hiding, chroma_weight, _p16(coeffs), res),
           "hb_tq_encode_intra")
    return coeffs, list(res)


Context.tq_encode_intra = _tq_encode_intra


def _intra_run(self, cur, pred, jobs, adi):
    """adi: int16 array, the jobs' 4*size+1 reference samples back to back.  Returns (n, 35) SADs (rows of prediction jobs are 0)"""
    n = len(jobs)
    arr = (IntraJob * n)(*jobs)
    sads = np.zeros((n, 35), np.uint32)
    _check(self.L.hb_intra_run(self.h, cur.h if cur is not None else None, pred.h if pred is not None else None, arr, n, _p16(adi),
                               sads.ctypes.data_as(C.POINTER(C.c_uint32))), "hb_intra_run")
    return sads


Context.intra_run = _intra_run

SAO_DT = np.dtype([("eo_diff", "<i4", (4, 5)), ("eo_count", "<i4", (4, 5)), ("bo_diff", "<i4", (32,)), ("bo_count", "<i4", (32,))])


def _sao_stats(self, orig, rec):
    """SAO statistics of every CTU and component: (n_ctus, 3) array of SAO_DT (hb_sao_stats)"""
    n = ((rec.w + 63) // 64) * ((rec.h_px + 63) // 64)
    out = np.zeros((n, 3), SAO_DT)
    _check(self.L.hb_sao_stats_frame(self.h, orig.h, rec.h, out.ctypes.data), "hb_sao_stats_frame")
    return out


Context.sao_stats = _sao_stats


def _deblock(self, frame, bs_ver, bs_hor, qp, cb_qp_offset=0, cr_qp_offset=0, beta_offset_div2=0, tc_offset_div2=0):
    """deblocking in place; bs_ver / bs_hor / qp: uint8 (rows, units_w) maps per 4x4 luma unit"""
    bs_ver = np.ascontiguousarray(bs_ver, np.uint8); bs_hor = np.ascontiguousarray(bs_hor, np.uint8); qp = np.ascontiguousarray(qp, np.uint8)
    assert bs_ver.shape == bs_hor.shape == qp.shape and bs_ver.shape[0] >= frame.h_px // 4
    prm = (C.c_int32 * 4)(cb_qp_offset, cr_qp_offset, beta_offset_div2, tc_offset_div2)
    _check(self.L.hb_deblock_frame(self.h, frame.h, bs_ver.ctypes.data, bs_hor.ctypes.data, qp.ctypes.data, bs_ver.shape[1], prm), "hb_deblock_frame")


Context.deblock = _deblock

UNIT_INFO_DT = np.dtype([("cu_depth", "u1"), ("tu_depth", "u1"), ("intra", "u1"), ("cbf_luma", "u1"), ("ref_idx", "i1"), ("qp", "u1"), ("mvx", "<i2"), ("mvy", "<i2")])


def _deblock_units(self, frame, units, cb_qp_offset=0, cr_qp_offset=0, beta_offset_div2=0, tc_offset_div2=0):
    """deblocking in place with the strengths derived on the device; units: (rows, units_w) array of UNIT_INFO_DT.  Returns (bs_ver, bs_hor)"""
    units = np.ascontiguousarray(units, UNIT_INFO_DT)
    bsv = np.zeros(units.shape, np.uint8); bsh = np.zeros(units.shape, np.uint8)
    prm = (C.c_int32 * 4)(cb_qp_offset, cr_qp_offset, beta_offset_div2, tc_offset_div2)
    _check(self.L.hb_deblock_frame_units(self.h, frame.h, units.ctypes.data, units.shape[1], prm, bsv.ctypes.data, bsh.ctypes.data), "hb_deblock_frame_units")
    return bsv, bsh


Context.deblock_units = _deblock_units

UNIT_L1_DT = np.dtype([("ref_idx", "i1"), ("reserved", "i1"), ("mvx", "<i2"), ("mvy", "<i2")])


def _deblock_units_b(self, frame, units, units_l1, pic_l0, pic_l1, cb_qp_offset=0, cr_qp_offset=0, beta_offset_div2=0, tc_offset_div2=0):
    """B pictures: units = list 0 (UNIT_INFO_DT), units_l1 = list 1 (UNIT_L1_DT), pic_l0 / pic_l1 = picture behind every reference index"""
    units = np.ascontiguousarray(units, UNIT_INFO_DT); units_l1 = np.ascontiguousarray(units_l1, UNIT_L1_DT)
    p0 = np.ascontiguousarray(pic_l0, np.int32); p1 = np.ascontiguousarray(pic_l1, np.int32)
    bsv = np.zeros(units.shape, np.uint8); bsh = np.zeros(units.shape, np.uint8)
    prm = (C.c_int32 * 4)(cb_qp_offset, cr_qp_offset, beta_offset_div2, tc_offset_div2)
    self.L.hb_deblock_frame_units_b.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(self.L.hb_deblock_frame_units_b(self.h, frame.h, units.ctypes.data, units_l1.ctypes.data, units.shape[1], p0.ctypes.data, len(p0), p1.ctypes.data, len(p1),
                                           prm, bsv.ctypes.data, bsh.ctypes.data), "hb_deblock_frame_units_b")
    return bsv, bsh


Context.deblock_units_b = _deblock_units_b


def _amvp_candidates(self, units, width, height, jobs):
    """AMVP predictors of 2Nx2N PUs from the per-unit motion field; units: (16 * CTU rows, units_w) UNIT_INFO_DT, jobs: (n, 3) int32
    (x, y, size).  Returns (n, 2, 2) int32: the two predictors (x, y) of every PU."""
    units = np.ascontiguousarray(units, UNIT_INFO_DT)
    jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 3)
    out = np.zeros((len(jobs), 2, 2), np.int32)
    _check(self.L.hb_amvp_candidates(self.h, units.ctypes.data, units.shape[1], width, height, jobs.ctypes.data, len(jobs), out.ctypes.data), "hb_amvp_candidates")
    return out


Context.amvp_candidates = _amvp_candidates

INTRA_UNIT_DT = np.dtype([(k, "<i4") for k in ("comp", "x", "y", "size", "mode", "qp", "scan_mode", "node_x", "node_y", "node_size")])


def _intra_reconstruct(self, cur, pred, recon, units, is_islice=1, sign_hiding=1, chroma_weight=1.0, per_level_launches=0):
    """wavefront-batched reconstruction of intra transform units; units: (n, 10) int32 or INTRA_UNIT_DT in coding order.
    Returns (levels back to back in unit order, results array {sum, ssd, ssd_zero, zeroed}, number of dependency levels)"""
    u = np.ascontiguousarray(units, np.int32).reshape(-1, 10)
    n = len(u)
    coeffs = np.zeros(int((u[:, 3].astype(np.int64) ** 2).sum()), np.int16)
    res = np.zeros(n, np.dtype([("sum", "<i4"), ("ssd", "<u4"), ("ssd_zero", "<u4"), ("zeroed", "<i4")]))
    levels = C.c_int32(0)
    self.L.hb_intra_reconstruct_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(self.L.hb_intra_reconstruct_ex(self.h, cur.h, pred.h, recon.h, u.ctypes.data, n, is_islice, sign_hiding, chroma_weight, 1 if per_level_launches else 0,
                                          coeffs.ctypes.data, res.ctypes.data, C.byref(levels)), "hb_intra_reconstruct")
    return coeffs, res, levels.value


Context.intra_reconstruct = _intra_reconstruct


def _merge_candidates(self, units, width, height, jobs, max_cands=5):
    """merge candidates of 2Nx2N PUs from the per-unit motion field: (n, max_cands, 2) int32"""
    units = np.ascontiguousarray(units, UNIT_INFO_DT)
    jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 3)
    out = np.zeros((len(jobs), max_cands, 2), np.int32)
    _check(self.L.hb_merge_candidates(self.h, units.ctypes.data, units.shape[1], width, height, jobs.ctypes.data, len(jobs), max_cands, out.ctypes.data), "hb_merge_candidates")
    return out


Context.merge_candidates = _merge_candidates

SAO_PARAM_DT = np.dtype([("type", "i1", (3,)), ("reserved", "i1"), ("offset", "<i2", (3, 32))])


def _sao_apply(self, src, dst, types, offsets):
    """SAO offset pass: types (n_ctus, 3) in -1..4, offsets (n_ctus, 3, 32) as sao_offset_t.offset holds them"""
    prm = np.zeros(len(types), SAO_PARAM_DT)
    prm["type"] = types; prm["offset"] = offsets
    _check(self.L.hb_sao_apply_frame(self.h, src.h, dst.h, prm.ctypes.data), "hb_sao_apply_frame")


Context.sao_apply = _sao_apply


def sao_derive_offsets(stats, sao_type, lam):
    """host arithmetic of the SAO decision on one (CTU, component) record of SAO_DT: (offsets[32] int16, band, dist)"""
    L = load_library()
    rec = np.ascontiguousarray(stats, SAO_DT).reshape(1)
    off = np.zeros(32, np.int16); band = C.c_int32(0); dist = C.c_int64(0)
    _check(L.hb_sao_derive_offsets(rec.ctypes.data, sao_type, float(lam), off.ctypes.data, C.byref(band), C.byref(dist)), "hb_sao_derive_offsets")
    return off, band.value, dist.value


def sao_decide_standin(stats, lambdas):
    """stand-in SAO decision over hb_sao_stats_frame's output (n_ctus, 3): returns the SAO_PARAM_DT array hb_sao_apply_frame takes"""
    L = load_library()
    stats = np.ascontiguousarray(stats, SAO_DT)
    prm = np.zeros(stats.shape[0], SAO_PARAM_DT)
    lam = (C.c_double * 3)(*[float(x) for x in lambdas])
    _check(L.hb_sao_decide_standin(stats.ctypes.data, stats.shape[0], lam, prm.ctypes.data), "hb_sao_decide_standin")
    return prm


SAO_CAND_DT = np.dtype([("dist", "<i8"), ("offset", "i1", (4,)), ("band", "i1"), ("reserved", "i1", (3,))])


def _sao_candidates(self, orig, rec, lambdas, want_stats=False):
    """SAO statistics + the offsets / band position / distortion estimate of all five types derived on the device: (n_ctus, 3, 5) SAO_CAND_DT
    (and the (n_ctus, 3) statistics when asked)"""
    n = ((rec.w + 63) // 64) * ((rec.h_px + 63) // 64)
    cand = np.zeros((n, 3, 5), SAO_CAND_DT)
    st = np.zeros((n, 3), SAO_DT) if want_stats else None
    lam = (C.c_double * 3)(*[float(x) for x in lambdas])
    _check(self.L.hb_sao_candidates_frame(self.h, orig.h, rec.h, lam, cand.ctypes.data, st.ctypes.data if want_stats else None), "hb_sao_candidates_frame")
    return (cand, st) if want_stats else cand


Context.sao_candidates = _sao_candidates


def sao_decide_from_candidates(cand, lambdas):
    L = load_library()
    cand = np.ascontiguousarray(cand, SAO_CAND_DT)
    prm = np.zeros(cand.shape[0], SAO_PARAM_DT)
    lam = (C.c_double * 3)(*[float(x) for x in lambdas])
    _check(L.hb_sao_decide_from_candidates(cand.ctypes.data, cand.shape[0], lam, prm.ctypes.data), "hb_sao_decide_from_candidates")
    return prm


def presearch_records(jobs_xyn):
    """int32 (n, 3) {x, y, size} -> the hb_intra_job records of an all-mode search (build once, reuse every frame of that size)"""
    rec = np.zeros((len(jobs_xyn), 6), np.int32)
    rec[:, 1:4] = jobs_xyn
    rec[:, 4] = -1; rec[:, 5] = -1                      # mode < 0: SADs of all modes; filtered < 0: the search rule
    return rec


def _intra_presearch(self, cur, jobs, adi, out=None):
    """all 35 mode SADs of many luma blocks at once.  jobs: presearch_records(...) (or the (n, 3) {x, y, size} array),
    adi: their reference samples back to back, out: optional (n, 35) uint32 table (pinned memory avoids a staging copy)."""
    rec = jobs if jobs.shape[1] == 6 else presearch_records(jobs)
    n = len(rec)
    arr = (IntraJob * n).from_buffer(rec)
    sads = out if out is not None else np.empty((n, 35), np.uint32)
    assert sads.dtype == np.uint32 and sads.size == n * 35 and adi.dtype == np.int16 and adi.flags.c_contiguous
    _check(self.L.hb_intra_run(self.h, cur.h, None, arr, n, _p16(adi), sads.ctypes.data_as(C.POINTER(C.c_uint32))), "hb_intra_run")
    return sads


Context.intra_presearch = _intra_presearch


class FrameIpc(C.Structure):
    _fields_ = [("mem", (C.c_uint8 * 64) * 3), ("width", C.c_int32), ("height", C.c_int32)]


class RowSpan(C.Structure):
    _fields_ = [("src", C.c_int32), ("plane", C.c_int32), ("row0", C.c_int32), ("n_rows", C.c_int32)]


class IpcEvent:
    """inter-process event on a context's stream: the owner records, another process's stream waits (no host wait)"""

    def __init__(self, ctx, handle=None):
        self.ctx = ctx
        ev = C.c_void_p()
        if handle is None:
            buf = (C.c_uint8 * 64)()
            _check(ctx.L.hb_ipc_event_create(ctx.h, C.byref(ev), buf), "hb_ipc_event_create")
            self.handle = bytes(buf)
        else:
            buf = (C.c_uint8 * 64).from_buffer_copy(handle)
            _check(ctx.L.hb_ipc_event_open(ctx.h, buf, C.byref(ev)), "hb_ipc_event_open")
            self.handle = bytes(handle)
        self.h = ev

    def record(self):
        _check(self.ctx.L.hb_ipc_event_record(self.ctx.h, self.h), "hb_ipc_event_record")

    def wait(self):
        _check(self.ctx.L.hb_ipc_event_wait(self.ctx.h, self.h), "hb_ipc_event_wait")

    def close(self):
        if self.h:
            self.ctx.L.hb_ipc_event_destroy(self.h)
            self.h = None


class Frame:
    def __init__(self, ctx, width, height):
        self.ctx, self.w, self.h_px = ctx, width, height
        h = C.c_void_p()
        _check(ctx.L.hb_frame_create(ctx.h, width, height, C.byref(h)), "hb_frame_create")
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.hb_frame_destroy(self.h)
            self.h = None

    def upload_u8(self, y, u, v):
        for a in (y, u, v):
            assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
        _check(self.ctx.L.hb_frame_upload_u8(self.ctx.h, self.h, y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0],
                                             v.ctypes.data, v.strides[0]), "hb_frame_upload_u8")

    def upload_i16(self, y, u, v):
        for a in (y, u, v):
            assert a.dtype == np.int16 and a.ndim == 2 and a.strides[1] == 2
        _check(self.ctx.L.hb_frame_upload_i16(self.ctx.h, self.h, y.ctypes.data, y.strides[0] // 2, u.ctypes.data,
                                              u.strides[0] // 2, v.ctypes.data, v.strides[0] // 2), "hb_frame_upload_i16")

    def export_rows(self, plane, row0, n_rows, dev_ptr):
        _check(self.ctx.L.hb_frame_export_rows(self.ctx.h, self.h, plane, row0, n_rows, dev_ptr), "hb_frame_export_rows")

    def import_rows(self, plane, row0, n_rows, dev_ptr):
        _check(self.ctx.L.hb_frame_import_rows(self.ctx.h, self.h, plane, row0, n_rows, dev_ptr), "hb_frame_import_rows")

    def pad(self):
        _check(self.ctx.L.hb_frame_pad(self.ctx.h, self.h), "hb_frame_pad")

    # ---- peer pictures (CUDA IPC): bytes another process opens with Frame.ipc_open; rows pulled straight out of the owner's HBM
    def ipc_export(self):
        d = FrameIpc()
        _check(self.ctx.L.hb_frame_ipc_export(self.h, C.byref(d)), "hb_frame_ipc_export")
        return bytes(d)

    @classmethod
    def ipc_open(cls, ctx, blob):
        d = FrameIpc.from_buffer_copy(blob)
        f = cls.__new__(cls)
        f.ctx, f.w, f.h_px = ctx, d.width, d.height
        h = C.c_void_p()
        _check(ctx.L.hb_frame_ipc_open(ctx.h, C.byref(d), C.byref(h)), "hb_frame_ipc_open")
        f.h = h
        return f

    def pull_rows_prepare(self, srcs, spans):
        return ((C.c_void_p * max(1, len(srcs)))(*[s.h for s in srcs]), len(srcs), (RowSpan * max(1, len(spans)))(*[RowSpan(*s) for s in spans]), len(spans))

    def pull_rows_prepared(self, prep, refresh_border=True):
        _check(self.ctx.L.hb_frame_pull_rows(self.ctx.h, self.h, prep[0], prep[1], prep[2], prep[3], 1 if refresh_border else 0), "hb_frame_pull_rows")

    def pull_rows(self, srcs, spans, refresh_border=True):
        """spans: [(index into srcs, plane, row0, n_rows)]: one copy kernel on the context's stream, nothing waits"""
        arr = (C.c_void_p * max(1, len(srcs)))(*[s.h for s in srcs])
        sp = (RowSpan * max(1, len(spans)))(*[RowSpan(*s) for s in spans])
        _check(self.ctx.L.hb_frame_pull_rows(self.ctx.h, self.h, arr, len(srcs), sp, len(spans), 1 if refresh_border else 0), "hb_frame_pull_rows")

    def download(self):
        y = np.zeros((self.h_px, self.w), np.uint8)
        u = np.zeros((self.h_px // 2, self.w // 2), np.uint8)
        v = np.zeros_like(u)
        _check(self.ctx.L.hb_frame_download_u8(self.ctx.h, self.h, y.ctypes.data, self.w, u.ctypes.data, self.w // 2,
                                               v.ctypes.data, self.w // 2), "hb_frame_download_u8")
        return y, u, v


class _BorrowedFrame(Frame):
    def __init__(self, ctx, handle, width, height):
        self.ctx, self.w, self.h_px, self.h = ctx, width, height, C.c_void_p(handle)

    def close(self):
        self.h = None


class Prepass:
    """frame-level pre-pass: every CTU, PU sizes 64..8, TU sizes 32..4 in one replay (include/homer_b200.h section D)"""
    DEPTHS, PASSES = 4, 5

    def __init__(self, ctx, width, height, qp=32, chroma_qp_offset=2, sign_hiding=1, is_islice=0, me_action=7,
                 use_graph=1, band=(0, 0), compact_tables=0, subpel_per_pu=0, me_staged_window=0, me_per_depth=0):
        self.ctx, self.w, self.h_px = ctx, width, height
        self.cfg = PrepassCfg(qp, chroma_qp_offset, sign_hiding, is_islice, me_action, use_graph, band[0], band[1], compact_tables, subpel_per_pu, me_staged_window, me_per_depth)
        h = C.c_void_p()
        _check(ctx.L.hb_prepass_create(ctx.h, width, height, C.byref(self.cfg), C.byref(h)), "hb_prepass_create")
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.hb_prepass_destroy(self.h)
            self.h = None

    def run(self, cur, ref, avg_dist):
        _check(self.ctx.L.hb_prepass_run(self.h, cur.h, ref.h, avg_dist), "hb_prepass_run")

    def run_profiled(self, cur, ref, avg_dist):
        """[(kernel name, device ms)] of one frame, launch by launch (no graph)"""
        ms = (C.c_float * 24)()
        n = self.ctx.L.hb_prepass_run_profiled(self.h, cur.h, ref.h, avg_dist, ms, 24)
        if n < 0:
            _check(n, "hb_prepass_run_profiled")
        return [(self.ctx.L.hb_prepass_kernel_name(self.h, i).decode(), ms[i]) for i in range(n)]

    def num_pus(self, depth):
        return self.ctx.L.hb_prepass_num_pus(self.h, depth)

    def num_tus(self, p, comp):
        return self.ctx.L.hb_prepass_num_tus(self.h, p, comp)

    def tu_size(self, p, comp):
        return self.ctx.L.hb_prepass_tu_size(self.h, p, comp)

    def tu_xy(self, p, comp):
        n = self.num_tus(p, comp)
        xy = np.zeros((n, 2), np.int32)
        if n:
            _check(self.ctx.L.hb_prepass_tu_xy(self.h, p, comp, xy.ctypes.data_as(i32p)), "hb_prepass_tu_xy")
        return xy

    def fetch_me(self, depth):
        n = self.num_pus(depth)
        out = (MeResult * n)()
        _check(self.ctx.L.hb_prepass_fetch_me(self.h, depth, out), "hb_prepass_fetch_me")
        return np.frombuffer(out, dtype=np.dtype([("mvx", "<i4"), ("mvy", "<i4"), ("subx", "<i4"), ("suby", "<i4"),
                                                  ("sad", "<u4"), ("n_probes", "<u4")])).copy()

    def fetch_tu(self, p, comp):
        n = self.num_tus(p, comp)
        out = (TuResult * max(n, 1))()
        if n:
            _check(self.ctx.L.hb_prepass_fetch_tu(self.h, p, comp, out), "hb_prepass_fetch_tu")
        return np.frombuffer(out, dtype=np.dtype([("sum", "<i4"), ("ssd", "<u4"), ("ssd_zero", "<u4"), ("zeroed", "<i4")]))[:n].copy()

    def fetch_coeffs(self, p, comp):
        n, t = self.num_tus(p, comp), self.tu_size(p, comp)
        out = np.zeros((n, t, t), np.int16)
        if n:
            _check(self.ctx.L.hb_prepass_fetch_coeffs(self.h, p, comp, _p16(out.reshape(-1))), "hb_prepass_fetch_coeffs")
        return out

    def pred(self, depth):
        return _BorrowedFrame(self.ctx, self.ctx.L.hb_prepass_pred(self.h, depth), self.w, self.h_px)

    def recon(self, p):
        return _BorrowedFrame(self.ctx, self.ctx.L.hb_prepass_recon(self.h, p), self.w, self.h_px)

    # ---- host-decision flow
    def tables_bytes(self):
        return int(self.ctx.L.hb_prepass_tables_bytes(self.h))

    def fetch_tables(self, pinned):
        """async copy of the cost tables (ME + TU records) into pinned memory; ctx.sync() before reading"""
        _check(self.ctx.L.hb_prepass_fetch_tables(self.h, pinned.ctypes.data, pinned.nbytes), "hb_prepass_fetch_tables")

    def num_ctus(self):
        return self.ctx.L.hb_prepass_num_ctus(self.h)

    def select(self, tables, lam, sel, ctu_off):
        _check(self.ctx.L.hb_prepass_select(self.h, tables.ctypes.data, lam, sel.ctypes.data, ctu_off.ctypes.data), "hb_prepass_select")

    def fetch_coeff_wnd(self, sel):
        """levels of the chosen passes in the reference's ctu->coeff_wnd layout: (n_ctus, 6144) int16 = 64*64 Y | 32*32 U | 32*32 V per CTU"""
        sel = np.ascontiguousarray(sel, np.uint8)
        out = np.zeros((self.num_ctus(), 64 * 64 + 2 * 32 * 32), np.int16)
        self.ctx.L.hb_prepass_fetch_coeff_wnd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _check(self.ctx.L.hb_prepass_fetch_coeff_wnd(self.h, sel.ctypes.data, out.ctypes.data), "hb_prepass_fetch_coeff_wnd")
        return out

    def gather(self, sel, ctu_off, pinned):
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_gather(self.h, sel.ctypes.data, ctu_off.ctypes.data, pinned.ctypes.data, pinned.nbytes, C.byref(n)),
               "hb_prepass_gather")
        return n.value

    def process_frame(self, cur, ref, cur_planes, ref_planes, avg_dist, lam, tables, sel, ctu_off, out):
        """upload -> pre-pass -> tables -> select -> gather as ONE blocking C call (releases the GIL for its whole duration)"""
        cp = (C.c_void_p * 3)(*[p.ctypes.data for p in cur_planes])
        rp = (C.c_void_p * 3)(*[p.ctypes.data for p in ref_planes])
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_process_frame(self.h, cur.h, ref.h, cp, rp, avg_dist, lam, tables.ctypes.data, tables.nbytes,
                                                   sel.ctypes.data, ctu_off.ctypes.data, out.ctypes.data, out.nbytes, C.byref(n)),
               "hb_prepass_process_frame")
        return n.value

    def frame_begin(self, cur, ref, cur_planes, ref_planes, avg_dist, tables):
        """first half of process_frame: queues uploads, pre-pass and table fetch, returns immediately"""
        cp = (C.c_void_p * 3)(*[p.ctypes.data for p in cur_planes])
        rp = (C.c_void_p * 3)(*[p.ctypes.data for p in ref_planes])
        _check(self.ctx.L.hb_prepass_frame_begin(self.h, cur.h, ref.h, cp, rp, avg_dist, tables.ctypes.data, tables.nbytes), "hb_prepass_frame_begin")

    def frame_finish(self, lam, tables, sel, ctu_off, out):
        """second half: waits for the tables, decides, gathers, waits for the results; returns the bytes written to `out`"""
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_frame_finish(self.h, lam, tables.ctypes.data, sel.ctypes.data, ctu_off.ctypes.data, out.ctypes.data, out.nbytes,
                                                  C.byref(n)), "hb_prepass_frame_finish")
        return n.value

    # ---- the same flow with the reference picture kept on the device
    def finalise(self, sel, ctu_off, rec, pinned_levels, cb_qp_offset=0, cr_qp_offset=0, beta_offset_div2=0, tc_offset_div2=0):
        """gather the choice into `rec`, deblock it there, queue the level streams to pinned memory; returns their byte count (async)"""
        prm = (C.c_int32 * 4)(cb_qp_offset, cr_qp_offset, beta_offset_div2, tc_offset_div2)
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_finalise(self.h, sel.ctypes.data, ctu_off.ctypes.data, rec.h, prm, pinned_levels.ctypes.data, pinned_levels.nbytes,
                                              C.byref(n)), "hb_prepass_finalise")
        return n.value

    def fetch_units(self):
        """(units, bs_ver, bs_hor) of the last finalise, each (height/4, width/4)"""
        shape = (self.h_px // 4, self.w // 4)
        units = np.zeros(shape, UNIT_INFO_DT); bsv = np.zeros(shape, np.uint8); bsh = np.zeros(shape, np.uint8)
        _check(self.ctx.L.hb_prepass_fetch_units(self.h, units.ctypes.data, bsv.ctypes.data, bsh.ctypes.data), "hb_prepass_fetch_units")
        return units, bsv, bsh

    def frame_begin_resident(self, cur, ref, cur_planes, avg_dist, tables):
        cp = (C.c_void_p * 3)(*[p.ctypes.data for p in cur_planes])
        _check(self.ctx.L.hb_prepass_frame_begin_resident(self.h, cur.h, ref.h, cp, avg_dist, tables.ctypes.data, tables.nbytes), "hb_prepass_frame_begin_resident")

    def frame_finish_resident(self, cur, lam, tables, sel, ctu_off, rec, next_ref, dbk, sao_lambda, levels, params=None):
        """dbk: four ints (cb, cr qp offsets, beta / tc offsets div 2); params (optional): (n_ctus,) SAO_PARAM_DT receiving the SAO decision.
        Returns the level bytes; next_ref is complete after ctx.sync() (or for anything queued on the same context)"""
        prm = (C.c_int32 * 4)(*dbk)
        lamv = (C.c_double * 3)(*[float(x) for x in sao_lambda])
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_frame_finish_resident(self.h, cur.h, lam, tables.ctypes.data, sel.ctypes.data, ctu_off.ctypes.data, rec.h, next_ref.h, prm, lamv,
                                                           levels.ctypes.data, levels.nbytes, C.byref(n), params.ctypes.data if params is not None else None),
               "hb_prepass_frame_finish_resident")
        return n.value

    def output_bytes(self):
        return int(self.ctx.L.hb_prepass_output_bytes(self.h))

    def fetch_all(self, pinned):
        n = C.c_size_t(0)
        _check(self.ctx.L.hb_prepass_fetch_all(self.h, pinned.ctypes.data, pinned.nbytes, C.byref(n)), "hb_prepass_fetch_all")
        return n.value
