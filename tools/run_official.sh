#!/bin/bash
# run_official.sh TAG -- the round's standard GPU pass (one B200): GPU tests, smoke, bench (both arms), launch list and one full ncu
# capture of the bench kernel.  Outputs under gpurun_out/<TAG>_*.
tag=${1:-r02}
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.txt; cat gpurun_out/${tag}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/${tag}_smoke.txt; cat gpurun_out/${tag}_smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 400 gpurun_out/${tag}_bench_ref.json
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --passes 1 --no-cpu > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pkz2048 -c 1 -f -o gpurun_out/${tag}_prof_pkz python bench.py --only-kernel --steps 2 --warmup 3 --passes 1 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
