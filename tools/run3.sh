mkdir -p gpurun_out
timeout 90 python tools/tc_smoke.py 2>&1 | tail -8 || { echo "TC SMOKE FAILED/HUNG - aborting"; exit 1; }
timeout 300 python tools/diag_conv.py 2>&1 | grep "^gen"
timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -12
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 124 -c 3 -o gpurun_out/prof_conv_r1 python tools/quick_time.py --res 512 --batch 16 --iters 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
