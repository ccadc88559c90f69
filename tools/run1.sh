set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_nontc.log; tail -30 gpurun_out/t_nontc.log
timeout 900 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/t_tc.log; tail -40 gpurun_out/t_tc.log
timeout 300 python tools/quick_time.py --res 256 --batch 8 --impl 0 2>&1 | tail -3
timeout 300 python tools/quick_time.py --res 512 --batch 4 --impl 0 2>&1 | tail -3
timeout 300 python tools/quick_time.py --res 512 --batch 4 --impl 0 --layers 2>&1 | tail -40
