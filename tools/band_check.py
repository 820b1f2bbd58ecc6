"""torchrun worker: CTU-row bands of one frame pair on WORLD_SIZE GPUs with an NCCL halo exchange of the reference frame.
Every rank uploads ONLY its own band of the reference (as if it had reconstructed it), swaps halos with its neighbours,
runs its band of the pre-pass and checks its PUs/TUs against a whole-frame run with the complete reference.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/band_check.py [W H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist

import homerhevc_b200 as hb
from homerhevc_b200 import bands, synth

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 720)
mode = sys.argv[3] if len(sys.argv) > 3 else "nccl"         # nccl: staged send/recv; peer: rows pulled out of the neighbours' HBM (CUDA IPC)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hb.Context(local)
tex = synth.make_texture(w, h)
cur_h, ref_h = synth.make_frame(tex, w, h, 3), synth.make_frame(tex, w, h, 2)
ctu_rows = (h + 63) // 64
row0, nrows = bands.band_ctu_rows(ctu_rows, world, rank)

# whole-frame truth on this GPU
cur, ref_full = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
cur.upload_u8(*cur_h); ref_full.upload_u8(*ref_h)
whole = hb.Prepass(ctx, w, h, use_graph=0)
whole.run(cur, ref_full, 650.0)

# banded: only this rank's rows of the reference are real, the rest is poison until the exchange fills the halos
masked = []
for c, pl in enumerate(ref_h):
    y0, y1 = bands.band_sample_rows(h, ctu_rows, world, rank, chroma=c > 0)
    m = np.full_like(pl, 255 if c == 0 else 7)
    m[y0:y1] = pl[y0:y1]
    masked.append(m)
ref_band = hb.Frame(ctx, w, h)
ref_band.upload_u8(*masked)
if mode == "peer":
    ctx.sync()
    ex = bands.PeerHaloPuller(hb, dist, ctx, [ref_band], w, h, world, rank)
    ex.mark_ready(0)
    ex.pull(0)
else:
    ex = bands.FrameHaloExchanger(torch, dist, ctx, w, h, world, rank, torch.device("cuda", local))
    ex.exchange(ref_band)
band = hb.Prepass(ctx, w, h, use_graph=0, band=(row0, nrows))
band.run(cur, ref_band, 650.0)
ctx.sync()

bad = 0
for d in range(4):
    s = 64 >> d
    gw = ((w + 63) // 64) * (64 // s)
    a, b = whole.fetch_me(d), band.fetch_me(d)
    for idx in range(len(a)):
        in_band = row0 <= (idx // gw) * s // 64 < row0 + nrows
        if in_band:
            bad += int(a[idx] != b[idx])
        else:
            bad += int(b[idx]["sad"] != 0xFFFFFFFF)
n_tu = 0
for p in range(5):
    for c in range(3):
        t = band.tu_size(p, c)
        if not t:
            continue
        xy_b, res_b, co_b = band.tu_xy(p, c), band.fetch_tu(p, c), band.fetch_coeffs(p, c)
        xy_w, res_w, co_w = whole.tu_xy(p, c), whole.fetch_tu(p, c), whole.fetch_coeffs(p, c)
        index = {(int(x), int(y)): i for i, (x, y) in enumerate(xy_w)}
        for i, (x, y) in enumerate(xy_b):
            j = index[(int(x), int(y))]
            bad += int(res_b[i] != res_w[j]) + int(not np.array_equal(co_b[i], co_w[j]))
            n_tu += 1
t = torch.tensor([bad, n_tu, ex.bytes_per_exchange], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print(f"BANDS world={world} {w}x{h} exchange={mode}: mismatches={int(t[0])} tus_checked={int(t[1])} halo_bytes_sent_total={int(t[2])}")
dist.barrier()                 # nobody closes a picture a neighbour may still be reading
dist.destroy_process_group()
sys.exit(1 if int(t[0]) else 0)
