"""Which members of the per-call GPU table make a whole encode differ from the reference's own table?  Encodes a clip once per
mask bit (1 sad/ssd16b, 2 predict/reconst, 4 interpolation, 8 transforms, 16 quant/inv_quant) and compares the streams.
usage: python tools/table_bisect.py WxHxN [force_intra] -> JSON line (also gpurun_out/table_bisect_<clip>.json)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np

import homerhevc_b200 as hb
from _encode import TableHook, describe_mismatch, encode, hook_addr, make_yuv

clip = sys.argv[1] if len(sys.argv) > 1 else "1280x720x1"
w, h, nf = (int(v) for v in clip.split("x"))
fi = int(sys.argv[2]) if len(sys.argv) > 2 else 0
masks = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8, 16, 31]
L = hb.load_library()
yuv = make_yuv(w, h, nf)
gold_bs, gold_rec, t_cpu = encode(w, h, yuv, nf, force_intra=fi)
out = {"clip": clip, "force_intra": fi, "cpu_seconds": round(t_cpu, 2), "masks": {}}
for m in masks:
    bs, rec, t = encode(w, h, yuv, nf, hook=hook_addr("refdrv_install_gpu_table"), user=TableHook(L._handle, m), force_intra=fi)
    same = bs == gold_bs and np.array_equal(rec, gold_rec)
    out["masks"][str(m)] = {"identical": bool(same), "seconds": round(t, 2), "what": None if same else describe_mismatch(w, h, bs, rec, gold_bs, gold_rec)}
    print(m, out["masks"][str(m)], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"table_bisect_{clip}.json"), "w") as f:
    f.write(json.dumps(out) + "\n")
print(json.dumps(out))
