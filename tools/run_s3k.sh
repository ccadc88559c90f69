mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider) 2>&1 | tail -3
timeout 200 python tools/quick_time.py --res 512 --batch 16 --iters 5 2>&1 | tail -1
timeout 200 python tools/quick_time.py --res 512 --batch 16 --layers --iters 3 > gpurun_out/s3_layers_final.txt 2>&1; tail -1 gpurun_out/s3_layers_final.txt
timeout 280 python bench.py --no-cpu-baseline > gpurun_out/s3_bench4.json 2> gpurun_out/s3_bench4.err; cut -c1-230 gpurun_out/s3_bench4.json
