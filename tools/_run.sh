set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_renderer.py tests/test_full_features.py -m gpu -x -q 2>&1 | tail -3
timeout 900 env MODES=wavefront python tools/ab.py libb200pt_base.so libb200pt.so 2>&1 | tail -3
timeout 300 python tools/stage_times.py 2>&1 | tail -5
