#!/bin/bash
# prof_cfgs.sh TAG -- one full ncu capture per BASELINE geometry (cfg1..cfg4 kernels), each after a plain run of the same command
# has exited 0; summaries: tools/ncu_summary.py / ncu_stalls.py / ncu_opmix.py -> profiles/<TAG>_*.txt
tag=${1:-prof}
prof() { # name regex geometry
  python bench.py --only-kernel --steps 2 --warmup 3 --passes 1 --geometry $3 > gpurun_out/${tag}_$1.plain.log 2>&1 || { echo "plain run failed: $1"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o gpurun_out/${tag}_$1 python bench.py --only-kernel --steps 2 --warmup 3 --passes 1 --geometry $3 > gpurun_out/${tag}_$1.log 2>&1
  tail -1 gpurun_out/${tag}_$1.log | cut -c1-200
}
prof cfg3_pk3 pk3_kernel 16384,4096,1
prof cfg4_pk2048 pk2048_kernel 2048,256,1
prof cfg1_pksmall pksmall_kernel 1024,512,1
prof cfg2_pkz pkz2048_kernel 2048,512,2
