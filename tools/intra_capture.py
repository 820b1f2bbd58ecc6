"""Worker: the reference's own encode of ONE synthetic picture (an I picture) with its decisions, levels and unfiltered reconstruction captured
CTU by CTU (oracle/ref_hooks.c: hmr_deblock_sao_pad_sync_ctu), written to an .npz.  Run in its own process with HB_REF_DIR pointing at the build
of the reference to use -- oracle/_ref/zinit by default: the as-is build's SSE4.2 intra predictors read automatic variables they never wrote
(oracle/Makefile), so its pictures depend on what the process left on the stack.
usage: python tools/intra_capture.py W H QP SIGN_HIDING SEED OUT.npz"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

from _encode import encode_and_capture, make_yuv

w, h, qp, sh, seed = (int(v) for v in sys.argv[1:6])
t0 = time.perf_counter()
a = encode_and_capture(w, h, make_yuv(w, h, 1, seed=seed), 1, qp=qp, sign_hiding=sh)
secs = time.perf_counter() - t0
np.savez(sys.argv[6], seconds=secs, slice_type=a["slice_type"], slice_qp=a["slice_qp"], recon_y=a["recon"][0], recon_u=a["recon"][1], recon_v=a["recon"][2], coeff=a["coeff"],
         **{k: a[k] for k in ("pred_depth", "part_size", "mode_y", "mode_c", "tr_idx", "qp", "pred_mode", "cbf_y", "cbf_u", "cbf_v")})
