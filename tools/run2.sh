mkdir -p gpurun_out
timeout 90 python tools/tc_smoke.py 2>&1 | tail -8 || { echo "TC SMOKE FAILED/HUNG - aborting"; exit 1; }
timeout 300 python tools/diag_conv.py 2>&1 | grep -v "only first64\|only rest\|with x=hi\|with w=hi\|both hi" | tail -20
timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -5
timeout 400 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -30
timeout 200 python tools/quick_time.py --res 256 --batch 8 --impl 0 2>&1 | tail -2
timeout 200 python tools/quick_time.py --res 512 --batch 4 --impl 0 2>&1 | tail -2
timeout 200 python tools/quick_time.py --res 512 --batch 16 --impl 0 2>&1 | tail -2
