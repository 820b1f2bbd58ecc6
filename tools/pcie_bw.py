"""Host<->device copy bandwidth on this box (pinned memory), alone and both directions at once: the ceiling of bench.py's e2e."""
import torch, time
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t) / 1e9
run(1, 1, 3)
print("H2D alone %.1f GB/s" % run(1, 0)); print("D2H alone %.1f GB/s" % run(0, 1)); print("both, each %.1f GB/s" % run(1, 1))
