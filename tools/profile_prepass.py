"""Small driver for ncu: the 1080p pre-pass (the workload of bench.py) on one stream, plain launches (no CUDA graph), so that
every kernel shows up once per frame in the launch list.  Usage under ncu: python tools/profile_prepass.py [frames] [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import homerhevc_b200 as hb
from homerhevc_b200 import synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w, h = {"720p": (1280, 720), "1080p": (1920, 1080), "2160p": (3840, 2160)}[sys.argv[2] if len(sys.argv) > 2 else "1080p"]
tex = synth.make_texture(w, h)
ctx = hb.Context(0)
res = []
for n in range(frames + 1):
    f = hb.Frame(ctx, w, h); f.upload_u8(*synth.make_frame(tex, w, h, n)); res.append(f)
pp = hb.Prepass(ctx, w, h, qp=32, use_graph=0, subpel_per_pu=int(os.environ.get("HB_SUBPEL_PER_PU", "0")), me_staged_window=int(os.environ.get("HB_STAGED_WINDOW", "0")), me_per_depth=int(os.environ.get("HB_ME_PER_DEPTH", "0")))      # 1: the round-1 per-PU plane kernels
for n in range(frames):
    pp.run(res[n + 1], res[n], 650.0)
ctx.sync()
print("done", frames, "frames", w, h)
