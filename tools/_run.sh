set -x
timeout 900 python tools/bench_configs.py > gpurun_out/r01d_configs.jsonl 2> gpurun_out/r01d_configs.err; cat gpurun_out/r01d_configs.jsonl | cut -c1-330; tail -3 gpurun_out/r01d_configs.err
