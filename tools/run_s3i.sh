mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 102 -c 16 -o gpurun_out/prof_conv_r1_final python tools/quick_time.py --res 512 --batch 16 --iters 1 --graphs 0 > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir4x4_2p\|fir4x4_nhwc -s 28 -c 4 -o gpurun_out/prof_fir_r1_final python tools/quick_time.py --res 512 --batch 16 --iters 1 --graphs 0 > gpurun_out/ncu_fir.log 2>&1; tail -1 gpurun_out/ncu_fir.log
