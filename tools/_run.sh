set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 3000 gpurun_out/bench_r01b.json; tail -5 gpurun_out/bench_r01b.err
timeout 200 env SPP=32 MODES=wavefront python tools/profile_run.py 2>&1 | tail -2 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_wf_extend' -s 22 -c 11 -f -o gpurun_out/r01b_prof_extend env SPP=32 MODES=wavefront python tools/profile_run.py 2>&1 | tail -3
