set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 env MODES=wavefront python tools/ab.py libb200pt_base.so libb200pt.so 2>&1 | tail -3
