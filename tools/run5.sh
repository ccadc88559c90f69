mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -5
timeout 200 python tools/tc_smoke.py 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -5
timeout 200 python tools/quick_time.py --res 512 --batch 16 --impl 0 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1b.csv python tools/quick_time.py --res 512 --batch 16 --iters 1 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
