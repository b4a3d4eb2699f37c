#!/bin/bash
# prof_cfg5.sh TAG [ENV=VAL] -- one full ncu capture of the cfg5 kernel (N = 65536, hop 1024, 1080 log rows)
tag=$1; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pkcl|pkcta2" -c 1 -f -o gpurun_out/${tag} python tools/bench_cfg5.py > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
