mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider) 2>&1 | tail -3
timeout 200 python tools/quick_time.py --res 512 --batch 16 --iters 5 2>&1 | tail -1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 20 -c 3 -o gpurun_out/prof_pair_s3 python tools/quick_time.py --res 512 --batch 16 --iters 1 --graphs 0 > gpurun_out/ncu_pair.log 2>&1; tail -1 gpurun_out/ncu_pair.log
