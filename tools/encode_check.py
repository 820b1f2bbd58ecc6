"""One whole-encode comparison in its own process: the reference's encoder in lock step as it is, and again with a replacement
installed -- `hooks` (CU-granularity GPU API + per-call GPU table, oracle/ref_hooks.c), `table` (per-call GPU table alone) or
`emu` (the hooks on the CPU emulation of the session, no GPU) -- and a byte-for-byte comparison of stream and reconstruction.
The reference build both arms run on is $HB_REF_DIR (default oracle/_ref; oracle/_ref/zinit = the same sources with
-ftrivial-auto-var-init=zero, see oracle/Makefile).  usage: python tools/encode_check.py WxHxN [perf] [mode] [batch] -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np

from _encode import CuHookCfg, TableHook, cu_hooks_off, describe_mismatch, encode, hook_addr, make_yuv

clip = sys.argv[1]
perf = int(sys.argv[2]) if len(sys.argv) > 2 else -1
mode = sys.argv[3] if len(sys.argv) > 3 else "hooks"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
w, h, nf = (int(v) for v in clip.split("x"))
yuv = make_yuv(w, h, nf)
gold_bs, gold_rec, t_cpu = encode(w, h, yuv, nf, perf=perf)
cnt = {}
if mode == "table":
    import homerhevc_b200 as hb
    bs, rec, t = encode(w, h, yuv, nf, hook=hook_addr("refdrv_install_gpu_table"), user=TableHook(hb.load_library()._handle, 31), perf=perf)
else:
    lib = None
    if mode == "hooks":
        import homerhevc_b200 as hb
        lib = hb.load_library()._handle
    try:
        bs, rec, t = encode(w, h, yuv, nf, hook=hook_addr("refdrv_install_cu_hooks"), user=CuHookCfg(lib, 31 if lib else 0, batch), perf=perf)
    finally:
        cnt = cu_hooks_off()
same = bs == gold_bs and np.array_equal(rec, gold_rec)
print(json.dumps({"clip": clip, "perf": perf, "mode": mode, "ref_dir": os.environ.get("HB_REF_DIR", "oracle/_ref"), "identical": bool(same), "bytes": len(gold_bs),
                  "what": None if same else describe_mismatch(w, h, bs, rec, gold_bs, gold_rec), "seconds_reference": round(t_cpu, 3), "seconds_replaced": round(t, 3),
                  "fps_reference": round(nf / t_cpu, 3), "fps_replaced": round(nf / t, 3), "hook_calls": cnt}))
