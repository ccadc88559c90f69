mkdir -p gpurun_out
SHGAN_FIR_VARIANT=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:fir4x4 -s 28 -c 4 -o gpurun_out/prof_fir_s3 python tools/quick_time.py --res 512 --batch 16 --iters 1 --graphs 0 > gpurun_out/ncu_fir.log 2>&1; tail -1 gpurun_out/ncu_fir.log
