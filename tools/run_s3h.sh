mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py c2 c4 c5 > gpurun_out/s3_configs.jsonl 2> gpurun_out/s3_configs.err; tail -3 gpurun_out/s3_configs.err; cut -c1-300 gpurun_out/s3_configs.jsonl
timeout 280 python bench.py > gpurun_out/s3_bench3.json 2> gpurun_out/s3_bench3.err; tail -2 gpurun_out/s3_bench3.err; cut -c1-200 gpurun_out/s3_bench3.json
