#!/bin/bash
# One GPU-box pass that refreshes what profiles/ holds for the current build (run under gpurun from the repo root):
#   1. the GPU test suite                                   -> gpurun_out/gputest.log
#   2. ncu metrics pass of one 32-sample render call        -> gpurun_out/r02_stats.csv   (tools/ncu_kernel_stats.py -> profiles/kernel_stats.json)
#   3. ncu --set full of one batch, every k_wf_ launch      -> gpurun_out/r02_wavefront.ncu-rep (tools/ncu_summary.py -> profiles/r02_wavefront_ncu.md)
#   4. launch list of the bench command                     -> gpurun_out/r02_launches_bench.csv
# Every profiled command first runs to completion without the profiler.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
SPP=32 MODES=wavefront python tools/profile_run.py > gpurun_out/profile_run.log 2>&1 || exit 1
cat gpurun_out/profile_run.log
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_bytes.sum,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/r02_stats.csv env SPP=32 MODES=wavefront python tools/profile_run.py > gpurun_out/ncu_stats.log 2>&1
SPP=32 FLIGHT=32 OPTS=wf_overlap=0 MODES=wavefront python tools/profile_run.py >> gpurun_out/profile_run.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 50 -c 25 -o gpurun_out/r02_wavefront -f \
    env SPP=32 FLIGHT=32 OPTS=wf_overlap=0 MODES=wavefront python tools/profile_run.py > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
