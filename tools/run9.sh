timeout 300 python -m pytest tests -m gpu -k "not tc" -q -p no:cacheprovider 2>&1 | tail -4
timeout 200 python tools/tc_smoke.py 2>&1 | tail -4
timeout 200 python tools/microbench.py fir 2>&1 | tail -6
timeout 200 python tools/quick_time.py --res 512 --batch 16 --impl 0 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -k "tc" -q -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -4
