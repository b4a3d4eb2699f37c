#!/bin/bash
# run_abl.sh OUT V1 V2 ... : device-resident cfg2-batch throughput of kernel-variant libraries tools/bin/variants/lib_<V>.so
# (an entry V@NAME=VALUE additionally sets one environment variable, e.g. z_base@JADE_PK_LOAD=pair2)
out=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=${v%%@*}; envs=""; [ "$lib" != "$v" ] && envs=${v#*@}
  echo -n "$v: " >> $out
  env $envs JADE_GPU_LIB=tools/bin/variants/lib_$lib.so timeout 300 python bench.py --only-kernel --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('%.1f Mframes/s  %.3f ms  frac %.4f  %s'%(d['value']/1e6, d.get('kernel_ms',0), d['value']*8196/1e9/6543.1, d.get('config',{}).get('kernel','')))
except Exception as e: print('ERR',e)" >> $out
done
cat $out
