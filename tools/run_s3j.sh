mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -p no:cacheprovider -k "pair" 2>&1 | tail -3
for impl in 4; do timeout 200 python tools/quick_time.py --res 512 --batch 16 --layers --iters 3 --impl $impl > gpurun_out/s3_layers_p${impl}b.txt 2>&1; tail -1 gpurun_out/s3_layers_p${impl}b.txt; done
