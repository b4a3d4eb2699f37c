echo "== cfg1 1024,512,1"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 1024,512,1 2>&1 | tail -1 | cut -c1-160
echo "== 1024,512,2"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 1024,512,2 2>&1 | tail -1 | cut -c1-160
echo "== 512,256,2"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 512,256,2 2>&1 | tail -1 | cut -c1-160
echo "== 128,64,1"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 128,64,1 2>&1 | tail -1 | cut -c1-160
echo "== cfg3 16384,4096,1"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 16384,4096,1 2>&1 | tail -1 | cut -c1-160
echo "== cfg4 2048,256,1"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 2048,256,1 2>&1 | tail -1 | cut -c1-160
echo "== 4096,1024,2"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 4096,1024,2 2>&1 | tail -1 | cut -c1-160
echo "== 8192,2048,2"; python bench.py --only-kernel --steps 10 --warmup 3 --geometry 8192,2048,2 2>&1 | tail -1 | cut -c1-160
python tools/bench_cfg5.py
