set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 env TILE=1 python tools/sweep_wf.py 2>&1 | tail -2
timeout 200 env SPP=32 MODES=wavefront python tools/profile_run.py 2>&1 | tail -2 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_wf_(primary|extend|bin|shade)' -s 68 -c 4 -f -o gpurun_out/r01b_prof_wavefront env SPP=32 MODES=wavefront python tools/profile_run.py 2>&1 | tail -3
