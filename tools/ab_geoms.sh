#!/bin/bash
# ab_geoms.sh OUT "GEOM ..." [REPEAT] -- device-resident throughput of the working-tree library against a baseline library
# (tools/bin/variants/lib_head.so, e.g. built from a `git worktree` of HEAD) on the given geometries N,hop,channels
# (bench.py --only-kernel).  The third-session A/B files gpurun_out/ab*.txt were made this way.
out=gpurun_out/${1:-ab.txt}; geoms=${2:-"2048,512,2 1024,512,1 16384,4096,1 2048,256,1"}; rep=${3:-1}; rm -f $out
run() { echo -n "$1 $2: " >> $out; env $3 timeout 300 python bench.py --only-kernel --steps 10 --warmup 3 --geometry $2 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('%.2f Mframes/s  %.4f ms'%(d['value']/1e6, d.get('kernel_ms',0)))" >> $out; }
for i in $(seq $rep); do for g in $geoms; do
  run head $g JADE_GPU_LIB=tools/bin/variants/lib_head.so
  run new  $g ""
done; done
cat $out
