"""Whole-encode parity at the drop-in boundary (north star: identical .265 bytes and reconstruction in fixed-QP mode,
WPP off).  The UNMODIFIED reference encoder (oracle/_ref) runs twice in lock step: once with its own
SSE4.2 function table, once with libhomer_b200's per-call drop-ins installed in that table exactly as INTEGRATION.md
section 1 shows -- every SAD / SSD / predict / reconst / interpolation / transform / quant / inv_quant call the host
mode decision makes then runs on the GPU.  The two bitstreams and reconstructions must be byte-identical."""
import ctypes as C

import numpy as np
import pytest

import homerhevc_b200 as hb
from homerhevc_b200 import synth
from _oracle import have_ref, ref

pytestmark = pytest.mark.gpu


class _Hook(C.Structure):
    _fields_ = [("lib", C.c_void_p), ("which", C.c_int)]


def _encode(D, w, h, yuv, nf, hook=None, user=None, force_intra=0, perf=-1):
    bs = np.zeros(4 << 20, np.uint8); rec = np.zeros(yuv.size, np.uint8); secs = C.c_double(0)
    n = D.refdrv_encode_lockstep(w, h, nf, yuv.ctypes.data_as(C.POINTER(C.c_uint8)), 32, 1, force_intra, perf,
                                 bs.ctypes.data_as(C.POINTER(C.c_uint8)), bs.size, rec.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 hook, user, C.byref(secs))
    assert n > 0, "reference encode failed"
    return bytes(bs[:n]), rec, secs.value


@pytest.mark.parametrize("which,nf,force_intra", [(31, 3, 0), (31, 2, 1)])
def test_bitstream_identical_with_gpu_table(ctx, which, nf, force_intra):
    if not have_ref():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    _, D = ref()
    L = hb.load_library()
    w, h = 192, 128
    clip = synth.make_clip(w, h, nf, seed=21)
    yuv = np.concatenate([np.concatenate([p.reshape(-1) for p in f]) for f in clip])
    gold_bs, gold_rec, t_cpu = _encode(D, w, h, yuv, nf, force_intra=force_intra)
    user = _Hook(L._handle, which)
    D.refdrv_install_gpu_table_addr.restype = C.c_void_p
    hook = C.c_void_p(D.refdrv_install_gpu_table_addr())
    gpu_bs, gpu_rec, t_gpu = _encode(D, w, h, yuv, nf, hook=hook, user=C.cast(C.pointer(user), C.c_void_p), force_intra=force_intra)
    assert len(gold_bs) > 200
    if gpu_bs != gold_bs or not np.array_equal(gpu_rec, gold_rec):
        # say as much as possible about a mismatch: where it starts, and whether a second run of either side repeats it
        first = next((i for i, (a, b) in enumerate(zip(gpu_bs, gold_bs)) if a != b), min(len(gpu_bs), len(gold_bs)))
        bad = np.flatnonzero(gpu_rec != gold_rec)
        frame = int(bad[0]) // (w * h * 3 // 2) if bad.size else -1
        gpu2, _, _ = _encode(D, w, h, yuv, nf, hook=hook, user=C.cast(C.pointer(user), C.c_void_p), force_intra=force_intra)
        gold2, _, _ = _encode(D, w, h, yuv, nf, force_intra=force_intra)
        pytest.fail(f"bitstreams differ from byte {first} ({len(gpu_bs)} vs {len(gold_bs)} bytes), {bad.size} reconstructed samples differ, first in frame {frame}; "
                    f"second GPU run {'equals' if gpu2 == gold_bs else 'differs from'} the CPU stream, second CPU run {'equals' if gold2 == gold_bs else 'differs from'} the first")
    D.refdrv_gpu_quant_calls.restype = C.c_long
    assert D.refdrv_gpu_quant_calls() > 100          # the GPU table really was on the path
    print(f"\nwhole encode {w}x{h}x{nf}: {len(gold_bs)} bytes identical; cpu {t_cpu:.2f}s, per-call gpu table {t_gpu:.2f}s")
