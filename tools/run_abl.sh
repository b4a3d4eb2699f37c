#!/bin/bash
# run_abl.sh OUT V1 V2 ... : device-resident cfg2-batch throughput of kernel-variant libraries tools/bin/variants/lib_<V>.so
out=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo -n "$v: " >> $out
  JADE_GPU_LIB=tools/bin/variants/lib_$v.so timeout 300 python bench.py --only-kernel --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('%.1f Mframes/s  %.3f ms  frac %.4f'%(d['value']/1e6, d.get('kernel_ms',0), d['value']*8196/1e9/6543.1))
except Exception as e: print('ERR',e)" >> $out
done
cat $out
