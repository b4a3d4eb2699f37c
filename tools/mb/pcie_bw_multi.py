"""Host <-> device copy bandwidth of pinned memory with N GPUs busy AT THE SAME TIME (one process per GPU under
torch.distributed.run): H2D alone, D2H alone and both directions at once, per GPU and summed over the box.  It is the ceiling of
bench.py's e2e leg (8194 bytes per frame over PCIe); whether that leg scales with the GPU count is a property of the host
(root complexes, IOMMU, memory bandwidth of the VM), which this prints.  Tooling only.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mb/pcie_bw_multi.py
"""
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 256 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=12):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([n * reps / dt / 1e9], dtype=torch.float64, device=dev)
    if world > 1:
        all_ = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(all_, gbs)
        return [float(x.item()) for x in all_]
    return [float(gbs.item())]


run(True, True, 2)
res = {"h2d_alone": run(True, False), "d2h_alone": run(False, True), "both_each_direction": run(True, True)}
if rank == 0:
    for k, v in res.items():
        print(f"N={world} {k:22s} per GPU min {min(v):6.1f} / mean {sum(v)/len(v):6.1f} / max {max(v):6.1f} GB/s   sum {sum(v):7.1f} GB/s"
              + ("  (x2 directions)" if k.startswith("both") else ""))
if world > 1:
    dist.destroy_process_group()
