mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider) 2>&1 | tail -4
timeout 200 python tools/quick_time.py --res 512 --batch 16 --iters 5 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_launches2.csv python tools/quick_time.py --res 512 --batch 16 --iters 1 --graphs 0 > gpurun_out/s3_ncu.log 2>&1
