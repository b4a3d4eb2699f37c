#!/bin/bash
# run_multi.sh N TAG -- N-GPU pass: parity on separate devices, concurrent PCIe ceiling, bench (both arms) under torchrun
n=$1; tag=$2
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4 > gpurun_out/${tag}_multi_tests.txt; cat gpurun_out/${tag}_multi_tests.txt
for g in 1 2 4 8; do
  [ $g -le $n ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 tools/mb/pcie_bw_multi.py 2>/dev/null | grep "^N=" >> gpurun_out/${tag}_pcie.txt
done
cat gpurun_out/${tag}_pcie.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -c 300 gpurun_out/${tag}_bench_n$n.json
for d in ${DEPTHS:-2 4 6}; do
  JADE_PIPE_DEPTH=$d python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --latency-blocks 200 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipe depth $d: e2e %.2f M frames/s, value %.1f M, shard_parity %s' % (d['e2e']['value']/1e6, d['value']/1e6, d.get('shard_parity')))" >> gpurun_out/${tag}_e2e_depth.txt
done
cat gpurun_out/${tag}_e2e_depth.txt
