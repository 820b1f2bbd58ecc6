"""Host-side logic that needs no GPU: synthetic clips, and the N>1 bench plumbing over gloo (world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from homerhevc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_clip_is_deterministic_and_pans():
    tex = synth.make_texture(256, 128, seed=3)
    y0, u0, v0 = synth.make_frame(tex, 256, 128, 0, noise=0.0)
    y1, _, _ = synth.make_frame(tex, 256, 128, 1, noise=0.0)
    assert y0.dtype == np.uint8 and y0.shape == (128, 256) and u0.shape == (64, 128)
    assert np.array_equal(y1[:-2, :-3], y0[2:, 3:])            # frame n is the texture shifted by (3n, 2n)
    assert np.array_equal(v0, 255 - u0)
    again = synth.make_frame(synth.make_texture(256, 128, seed=3), 256, 128, 0, noise=0.0)[0]
    assert np.array_equal(again, y0)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the same reduction bench.py applies to its per-rank timings: MAX over ranks, value = world * steps / max
t = torch.tensor([10.0 + 5.0 * rank, 20.0 - rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    print("MAX", t.tolist(), "VALUE", world * 100 / (t[0].item() * 1e-3))
dist.destroy_process_group()
'''


def test_multi_rank_timing_reduction_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAX [15.0, 20.0]" in out.stdout and "VALUE 13333.3" in out.stdout
