"""Whole-stream byte identity with the BATCHED API inside the reference's own encoder loop (north star: identical .265 bytes and
reconstruction in fixed-QP mode, WPP off).  The unmodified reference (oracle/_ref) runs in lock step twice: as it is, and with
oracle/ref_hooks.c interposing hmr_motion_estimation -> hb_enc_me (real AMVP lists and start points), hmr_motion_compensation_* ->
hb_enc_predict, encode_inter_cu(_chroma) -> hb_enc_tq (include/homer_b200.h section E) while everything else the host loop calls
through its function table (intra pictures, intra units of P pictures) runs on the per-call GPU table.  Sizes: BASELINE.json's
1280x720 (whose last CTU row is 16 samples high), IPPP, >= 10 pictures, in the default performance mode and in mode 0 (the only
one that searches 64x64 units), plus a picture with a partial CTU column."""
import ctypes as C

import numpy as np
import pytest

import homerhevc_b200 as hb
from _encode import CuHookCfg, cu_hooks_off, describe_mismatch, encode, hook_addr, make_yuv
from _oracle import have_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,nf,perf,batch", [(1280, 720, 10, -1, 1), (1280, 720, 4, 0, 1), (328, 200, 5, 0, 0), (192, 128, 3, -1, 1)])
def test_stream_identical_with_batched_api_in_the_encoder_loop(ctx, w, h, nf, perf, batch):
    if not have_ref():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    L = hb.load_library()
    yuv = make_yuv(w, h, nf)
    gold_bs, gold_rec, t_cpu = encode(w, h, yuv, nf, perf=perf)
    cfg = CuHookCfg(L._handle, 31, batch)
    try:
        bs, rec, t_gpu = encode(w, h, yuv, nf, hook=hook_addr("refdrv_install_cu_hooks"), user=cfg, perf=perf)
    finally:
        cnt = cu_hooks_off()
    assert bs == gold_bs and np.array_equal(rec, gold_rec), describe_mismatch(w, h, bs, rec, gold_bs, gold_rec) + f" {cnt}"
    assert cnt["frames"] == nf and cnt["p_frames"] == nf - 1
    assert cnt["me"] > 50 and cnt["mc"] > cnt["me"] and cnt["tq"] > 100
    assert cnt["me_fwd"] == cnt["mc_fwd"] == cnt["tq_fwd"] == cnt["tq_stale"] == cnt["errors"] == 0, cnt
    print(f"\nwhole encode {w}x{h}x{nf} perf {perf}: {len(gold_bs)} bytes identical; cpu {t_cpu:.2f}s ({nf / t_cpu:.2f} f/s), "
          f"batched API in the loop {t_gpu:.2f}s ({nf / t_gpu:.2f} f/s); hook calls {cnt}")
