mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 102 -c 7 -o gpurun_out/prof_halo_s3b python tools/quick_time.py --res 512 --batch 16 --iters 1 --impl 3 --graphs 0 > gpurun_out/ncu_halo.log 2>&1; tail -1 gpurun_out/ncu_halo.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 102 -c 7 -o gpurun_out/prof_tap_s3b python tools/quick_time.py --res 512 --batch 16 --iters 1 --impl 2 --graphs 0 > gpurun_out/ncu_tap.log 2>&1; tail -1 gpurun_out/ncu_tap.log
