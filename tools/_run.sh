set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01d_n1.json 2> gpurun_out/bench_r01d_n1.err; wc -l gpurun_out/bench_r01d_n1.json; tail -c 1200 gpurun_out/bench_r01d_n1.json
timeout 600 python bench.py --steps 3 --warmup 3 --mode megakernel --no-cpu-baseline > gpurun_out/bench_r01d_n1_mega.json 2>> gpurun_out/bench_r01d_n1.err; head -c 400 gpurun_out/bench_r01d_n1_mega.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
