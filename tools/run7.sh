mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/microbench.py fir 2>&1 | tail -8
timeout 200 python tools/quick_time.py --res 512 --batch 16 --impl 0 2>&1 | tail -1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fir4x4 -c 2 -o gpurun_out/prof_fir_r1c python tools/microbench.py fir > gpurun_out/ncu_fir.log 2>&1; tail -1 gpurun_out/ncu_fir.log
