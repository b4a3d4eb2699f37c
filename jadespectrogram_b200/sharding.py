"""Host-side sharding of a batch render over ranks / GPUs: no data-path collective.

The partition rules are the ones `jade_render_batch_multi` (csrc/jade_gpu.cu) applies inside one process; this module
states them once for the one-process-per-GPU launch (`bench.py` under torchrun, offline drivers) and for the CPU tests,
where the oracle stands in for the device.  Nothing here computes a spectrogram.

  * many streams  -> contiguous stream ranges per rank (BASELINE config 4);
  * one stream    -> contiguous column ranges per rank, each reading its own input range incl. the N-hop halo that it
                     shares with its neighbour (re-read, never exchanged; configs 3 and 5);
  * results       -> every rank owns a disjoint slab of the output; gathering is a concatenation.
Timing of a multi-rank job is the MAX over ranks of the device time (`reduce_max`).
"""
from dataclasses import dataclass


def partition(n_items, world, rank):
    """[lo, hi) of `n_items` for `rank` of `world` (same integer rule as jade_render_batch_multi)."""
    return n_items * rank // world, n_items * (rank + 1) // world


@dataclass
class Shard:
    stream_lo: int
    stream_hi: int
    col_lo: int
    col_hi: int

    @property
    def empty(self):
        return self.stream_hi <= self.stream_lo or self.col_hi <= self.col_lo


def plan(nstreams, ncols, world, rank):
    """The shard of an [nstreams][ncols] job that `rank` renders."""
    if nstreams >= world:
        lo, hi = partition(nstreams, world, rank)
        return Shard(lo, hi, 0, ncols)
    if nstreams == 1:
        lo, hi = partition(ncols, world, rank)
        return Shard(0, 1, lo, hi)
    return Shard(rank, rank + 1, 0, ncols) if rank < nstreams else Shard(0, 0, 0, 0)


def frame_start(j, hop, frames_per_block, block_stride, preroll):
    """Absolute index of the first sample of column j (jade_config geometry; Spectrogram.cpp:50-56 generalised)."""
    return (j // frames_per_block) * block_stride + (j % frames_per_block) * hop - preroll


def input_range(col_lo, col_hi, nsamples, fft_size, hop, frames_per_block=1, block_stride=None, preroll=None):
    """Sample range [a, b) that columns [col_lo, col_hi) read -- the shard's input incl. its halo."""
    block_stride = hop * frames_per_block if block_stride is None else block_stride
    preroll = fft_size if preroll is None else preroll
    a = max(0, frame_start(col_lo, hop, frames_per_block, block_stride, preroll))
    b = min(nsamples, frame_start(col_hi - 1, hop, frames_per_block, block_stride, preroll) + fft_size)
    return a, max(a, b)


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11} (sysfs cpulist format)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu_node(pci_domain, pci_bus, pci_device, sysfs="/sys/bus/pci/devices"):
    """One process per GPU: run this process (and so first-touch its pinned host buffers, the H2D source and D2H
    destination of `jade_render_batch`) on the CPUs of the NUMA node the GPU's PCIe root hangs off, so that DMA does not
    cross the socket interconnect.  Does nothing when the machine reports no NUMA node for the device (single-socket
    boxes, VMs without a virtual topology) or when the node's CPUs are not in the current affinity mask.
    Returns the NUMA node bound to, or None."""
    import os
    base = f"{sysfs}/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0"
    try:
        node = int(open(f"{base}/numa_node").read())
        local = parse_cpulist(open(f"{base}/local_cpulist").read())
    except (OSError, ValueError):
        return None
    if node < 0 or not hasattr(os, "sched_getaffinity"):
        return None
    allowed = os.sched_getaffinity(0)
    target = local & allowed
    if not target or target == allowed:
        return None
    os.sched_setaffinity(0, target)
    return node


def reduce_max(value, device=None):
    """MAX over ranks of a python float (device time of the slowest rank); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_slabs(local, axis_sizes, dst=0):
    """Gather per-rank numpy slabs (same trailing shape, different leading length) on rank `dst`; returns the list there."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    out = [None] * world
    dist.gather_object(local, out if dist.get_rank() == dst else None, dst=dst)
    return out if dist.get_rank() == dst else None
