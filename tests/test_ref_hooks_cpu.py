"""The reference-side binding of the CU-granularity API (oracle/ref_hooks.c) validated WITHOUT a GPU: the reference's own encoder
loop runs with hmr_motion_estimation / hmr_motion_compensation_* / encode_inter_cu(_chroma) interposed, and the hooks' jobs are
computed by an emulation of the session semantics of include/homer_b200.h section E built on the oracle's restatement.  The
streams must equal the unmodified reference's byte for byte -- which pins the argument mapping, the frame capture at hmr_rd_init,
the prediction-picture tracking and the result write-back of the hooks; the GPU tests then only swap the back end."""
import numpy as np
import pytest

from _encode import CuHookCfg, cu_hooks_off, describe_mismatch, encode, hook_addr, make_yuv
from _oracle import have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref was not built (needs /root/reference at build time)")


@pytest.mark.parametrize("w,h,nf,perf,batch", [(192, 128, 3, -1, 1), (328, 200, 4, 0, 1), (256, 144, 3, -1, 0)])
def test_cu_hooks_with_the_emulated_session_give_identical_streams(w, h, nf, perf, batch):
    yuv = make_yuv(w, h, nf)
    gold_bs, gold_rec, _ = encode(w, h, yuv, nf, perf=perf)
    cfg = CuHookCfg(None, 0, batch)
    try:
        bs, rec, _ = encode(w, h, yuv, nf, hook=hook_addr("refdrv_install_cu_hooks"), user=cfg, perf=perf)
    finally:
        cnt = cu_hooks_off()
    assert bs == gold_bs and np.array_equal(rec, gold_rec), describe_mismatch(w, h, bs, rec, gold_bs, gold_rec)
    # the hooks really were on the path of every P picture, and nothing had to be handed back to the host functions
    assert cnt["frames"] == nf and cnt["p_frames"] == nf - 1
    assert cnt["me"] > 50 and cnt["mc"] > cnt["me"] and cnt["tq"] > 100
    assert cnt["me_fwd"] == cnt["mc_fwd"] == cnt["tq_fwd"] == cnt["tq_stale"] == cnt["errors"] == 0, cnt
    assert cnt["mc_cached"] == 2 * cnt["mc"]                               # both chroma planes came with their luma call
    if batch:
        assert cnt["tq_cached"] == 2 * cnt["tq"]                           # and both chroma T/Q calls with their luma call


def test_inactive_hooks_forward_to_the_reference():
    """with no hooks installed the interposed symbols must be transparent: two plain encodes agree and count nothing"""
    w, h, nf = 128, 64, 2
    yuv = make_yuv(w, h, nf, seed=5)
    a, ra, _ = encode(w, h, yuv, nf)
    b, rb, _ = encode(w, h, yuv, nf)
    assert a == b and np.array_equal(ra, rb) and len(a) > 100
    cnt = cu_hooks_off()
    assert cnt["frames"] == 0 and cnt["me"] == 0


def test_cu_hooks_emulated_on_the_zero_initialised_build_at_720p():
    """the same check at BASELINE.json's 1280x720 (partial last CTU row), both arms on oracle/_ref/zinit (see tests/test_gpu_encode_hooks.py
    for why), in its own process"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    zinit = os.path.join(root, "oracle", "_ref", "zinit")
    if not os.path.exists(os.path.join(zinit, "librefdrv.so")):
        pytest.skip("oracle/_ref/zinit was not built")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "encode_check.py"), "1280x720x3", "-1", "emu", "1"], capture_output=True, text=True,
                         timeout=900, cwd=root, env=dict(os.environ, HB_REF_DIR=zinit))
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["identical"], r
    assert r["hook_calls"]["me"] > 1000 and r["hook_calls"]["tq_fwd"] == 0 and r["hook_calls"]["errors"] == 0
