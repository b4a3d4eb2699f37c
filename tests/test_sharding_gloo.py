"""The N > 1 path on CPU: two ranks over torch.distributed `gloo`, the oracle standing in for the device.

Checks that the partition rules of jadespectrogram_b200.sharding (the ones jade_render_batch_multi applies on real GPUs)
produce, rank by rank, slabs whose concatenation is bit-identical to the single-process rendering -- for the
stream-sharded (BASELINE config 4) and the column-sharded single stream with its N-hop halo (configs 3, 5) -- and that
the job time is the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle_lib as O
import signals

FS = 48000.0
pytestmark = pytest.mark.timeout(300)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, q):
    import torch.distributed as dist

    from jadespectrogram_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, hop, ch, S, n = case
        x = signals.streams(S, ch, n, FS, kind="mix", seed=77)  # every rank can regenerate its inputs (synthetic)
        ncols = n // hop + 1
        sh = sharding.plan(S, ncols, world, rank)
        slabs = []
        for s in range(sh.stream_lo, sh.stream_hi):
            # the shard reads only its own input range (incl. the halo), re-based like jade_render_batch does
            a, b = sharding.input_range(sh.col_lo, sh.col_hi, n, N, hop)
            assert a % hop == 0
            first = sh.col_lo - a // hop
            db, pix = O.render_batch(x[s][:, a:b], fft_size=N, hop=hop, first_col=first, ncols=sh.col_hi - sh.col_lo)
            slabs.append((db, pix))
        t_job = sharding.reduce_max(1.0 + rank)  # the slowest rank defines the job time
        got = sharding.gather_slabs((sh, slabs), None)
        if rank == 0:
            q.put((t_job, got))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(512, 128, 2, 5, 128 * 40), (1024, 256, 1, 1, 256 * 61), (256, 64, 1, 3, 64 * 30)])
def test_two_rank_sharding_reproduces_single_process(case):
    N, hop, ch, S, n = case
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_job, got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_job == 2.0
    x = signals.streams(S, ch, n, FS, kind="mix", seed=77)
    ncols = n // hop + 1
    covered = np.zeros((S, ncols), bool)
    for sh, slabs in got:
        for i, s in enumerate(range(sh.stream_lo, sh.stream_hi)):
            db, pix = slabs[i]
            odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, first_col=sh.col_lo, ncols=sh.col_hi - sh.col_lo)
            assert np.array_equal(db.view(np.uint32), odb.view(np.uint32)) and np.array_equal(pix, opix)
            assert not covered[s, sh.col_lo:sh.col_hi].any(), "shards overlap"
            covered[s, sh.col_lo:sh.col_hi] = True
    assert covered.all(), "shards do not cover the job"


def test_partition_rules():
    from jadespectrogram_b200 import sharding
    for n in (1, 2, 7, 1024, 11251):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.partition(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
    # halo: a column shard reads N - hop samples that its left neighbour also reads
    a0, b0 = sharding.input_range(0, 100, 10**9, 2048, 512)
    a1, b1 = sharding.input_range(100, 200, 10**9, 2048, 512)
    assert (a0, b0) == (0, 99 * 512) and a1 == 100 * 512 - 2048 and b0 - a1 == 2048 - 512
    # three streams on eight ranks: one stream each, the rest idle
    assert [sharding.plan(3, 50, 8, r).empty for r in range(8)] == [False] * 3 + [True] * 5


def test_bind_host_to_gpu_node_reads_sysfs(tmp_path):
    """NUMA binding helper: parses the sysfs cpulist, ignores devices without a node, never widens the affinity mask."""
    import os
    from jadespectrogram_b200 import sharding
    assert sharding.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    dev = tmp_path / "0000:d1:00.0"
    dev.mkdir()
    (dev / "numa_node").write_text("-1\n")
    (dev / "local_cpulist").write_text("0-1\n")
    before = os.sched_getaffinity(0)
    assert sharding.bind_host_to_gpu_node(0, 0xD1, 0, sysfs=str(tmp_path)) is None      # no NUMA node reported
    assert sharding.bind_host_to_gpu_node(0, 0x17, 0, sysfs=str(tmp_path)) is None      # no such device
    assert os.sched_getaffinity(0) == before
    (dev / "numa_node").write_text("1\n")
    (dev / "local_cpulist").write_text(",".join(str(c) for c in sorted(before)) + "\n")
    assert sharding.bind_host_to_gpu_node(0, 0xD1, 0, sysfs=str(tmp_path)) is None      # node covers every allowed CPU
    if len(before) > 1:
        one = min(before)
        (dev / "local_cpulist").write_text(f"{one},100000\n")
        try:
            assert sharding.bind_host_to_gpu_node(0, 0xD1, 0, sysfs=str(tmp_path)) == 1
            assert os.sched_getaffinity(0) == {one}
        finally:
            os.sched_setaffinity(0, before)
