"""Aggregate host<->device bandwidth with one process per GPU copying in both directions at once (pinned memory, 32 MiB chunks).
torchrun --nproc-per-node N tools/pcie_bw_multi.py   -- explains the e2e ceiling of bench.py at N > 1."""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 32 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter(); iters = 40
        for _ in range(iters):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    gbs = torch.tensor([n * iters / dt / 1e9], device="cuda")
    if world > 1:
        dist.all_reduce(gbs)
    if local == 0:
        print("%s: %.1f GB/s per direction summed over %d GPUs (%.1f per GPU)" % (mode, gbs.item(), world, gbs.item() / world), flush=True)
if world > 1:
    dist.destroy_process_group()
