mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv.py -m gpu -x -q -p no:cacheprovider -k "not tap and not halo" 2>&1 | tail -4
timeout 200 python tools/quick_time.py --res 512 --batch 16 --layers --iters 3 > gpurun_out/s3_layers_fir2.txt 2>&1; grep -E "fir|total" gpurun_out/s3_layers_fir2.txt | head -16
timeout 200 python tools/quick_time.py --res 512 --batch 16 --iters 5 2>&1 | tail -1
