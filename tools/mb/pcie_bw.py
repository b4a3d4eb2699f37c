"""PCIe ceiling for the e2e leg: pinned H2D and D2H copies alone and concurrently (two streams). Tooling only."""
import torch, time
n = 512 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 2)
print(f"H2D alone {run(True, False):.1f} GB/s; D2H alone {run(False, True):.1f} GB/s; concurrent: {run(True, True):.1f} GB/s each direction")
