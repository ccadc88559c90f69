timeout 300 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -4
for b in 1 4 16; do timeout 200 python tools/quick_time.py --res 512 --batch $b --graphs 0 2>&1 | tail -1; timeout 200 python tools/quick_time.py --res 512 --batch $b --graphs 1 2>&1 | tail -1; done
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_n1c.json | cut -c1-2500
