mkdir -p gpurun_out
timeout 200 python tools/microbench.py 2>&1 | tail -20
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fir4x4 -c 2 -o gpurun_out/prof_fir_r1b python tools/microbench.py fir > gpurun_out/ncu_fir.log 2>&1; tail -1 gpurun_out/ncu_fir.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 3 -c 2 -o gpurun_out/prof_dense_r1 python tools/microbench.py dense > gpurun_out/ncu_dense.log 2>&1; tail -1 gpurun_out/ncu_dense.log
