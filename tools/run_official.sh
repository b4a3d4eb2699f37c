set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 300 gpurun_out/bench_r2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 300 gpurun_out/bench_r2_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list_r2.log 2>&1
bash tools/run_cfgs.sh > gpurun_out/cfgs5.txt 2>&1
