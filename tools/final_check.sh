mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider) 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 280 python bench.py > gpurun_out/r1_bench_n1_final.json 2> gpurun_out/r1_bench_n1_final.err; cut -c1-200 gpurun_out/r1_bench_n1_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
