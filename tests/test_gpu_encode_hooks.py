"""Whole-stream byte identity with the BATCHED API inside the reference's own encoder loop (north star: identical .265 bytes and
reconstruction in fixed-QP mode, WPP off).  The reference's encoder runs in lock step twice: as it is, and with oracle/ref_hooks.c
interposing hmr_motion_estimation -> hb_enc_me (real AMVP lists and start points), hmr_motion_compensation_* -> hb_enc_predict,
encode_inter_cu(_chroma) -> hb_enc_tq (include/homer_b200.h section E) while everything else the host loop calls through its
function table (intra pictures, intra units of P pictures) runs on the per-call GPU table.  Sizes: BASELINE.json's 1280x720 (whose
last CTU row is 16 samples high), IPPP, >= 10 pictures, in the default performance mode and in mode 0 (the only one that searches
64x64 units), plus a picture with a partial CTU column.

Which build of the reference: its SSE4.2 intra predictors read automatic variables they never wrote, so the stream of the as-is
build depends on what earlier calls left on the stack (at 1280x720 the as-is / zero- / pattern-initialised builds give three
different streams, tools/ref_uninit_probe.sh) -- no replacement table can reproduce that.  Both arms of these tests therefore run on
oracle/_ref/zinit, the same unmodified sources compiled with -ftrivial-auto-var-init=zero (oracle/Makefile), each case in its own
process (two builds of the library must not meet in one).  The small as-is case of tests/test_gpu_whole_encode.py stays."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZINIT = os.path.join(ROOT, "oracle", "_ref", "zinit")
pytestmark = pytest.mark.gpu


def _check(clip, perf, mode, batch=1, ref_dir=ZINIT):
    if not os.path.exists(os.path.join(ref_dir, "librefdrv.so")):
        pytest.skip(f"{ref_dir} was not built (needs /root/reference at build time)")
    env = dict(os.environ, HB_REF_DIR=ref_dir)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "encode_check.py"), clip, str(perf), mode, str(batch)], capture_output=True, text=True,
                         timeout=1500, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print("\n" + json.dumps(r))
    return r


@pytest.mark.parametrize("clip,perf,batch", [("1280x720x10", -1, 1), ("1280x720x4", 0, 1), ("328x200x5", 0, 0), ("192x128x3", -1, 1)])
def test_stream_identical_with_batched_api_in_the_encoder_loop(clip, perf, batch):
    r = _check(clip, perf, "hooks", batch)
    nf = int(clip.split("x")[2])
    cnt = r["hook_calls"]
    assert r["identical"], f"{r['what']} {cnt}"
    assert cnt["frames"] == nf and cnt["p_frames"] == nf - 1
    assert cnt["me"] > 50 and cnt["mc"] > cnt["me"] and cnt["tq"] > 100
    assert cnt["me_fwd"] == cnt["mc_fwd"] == cnt["tq_fwd"] == cnt["tq_stale"] == cnt["errors"] == 0, cnt


def test_stream_identical_with_per_call_table_at_720p():
    """the literal drop-in (one launch per table call) on a 1280x720 intra + inter pair: the size at which the as-is build's dependence on
    stack contents first showed"""
    r = _check("1280x720x2", -1, "table")
    assert r["identical"], r["what"]
