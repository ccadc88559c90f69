mkdir -p gpurun_out
SHGAN_FIR_F32_2P=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv.py -m gpu -x -q -p no:cacheprovider -k "not tap and not halo" 2>&1 | tail -3
SHGAN_FIR_F32_2P=1 timeout 200 python tools/quick_time.py --res 512 --batch 16 --layers --iters 3 > gpurun_out/s3_layers_fir3.txt 2>&1; grep -E "fir|total" gpurun_out/s3_layers_fir3.txt | head -8
timeout 200 python tools/quick_time.py --res 512 --batch 16 --layers --iters 3 > gpurun_out/s3_layers_fir4.txt 2>&1; grep -E "fir|total" gpurun_out/s3_layers_fir4.txt | head -8
