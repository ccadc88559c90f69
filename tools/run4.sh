mkdir -p gpurun_out
timeout 90 python tools/tc_smoke.py 2>&1 | tail -4 || { echo "TC SMOKE FAILED/HUNG - aborting"; exit 1; }
timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/quick_time.py --res 512 --batch 16 --impl 0 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 102 -c 7 -o gpurun_out/prof_conv_r1b python tools/quick_time.py --res 512 --batch 16 --iters 2 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir4x4 -s 28 -c 3 -o gpurun_out/prof_fir_r1 python tools/quick_time.py --res 512 --batch 16 --iters 2 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
