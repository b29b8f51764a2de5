set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 env MODES=wavefront python tools/ab.py libb200pt_base.so libb200pt.so 2>&1 | tail -3
timeout 300 python tools/stage_times.py 2>&1 | sed -n 1,2p
timeout 300 env OPTS=wf_rel_primary=0 python tools/stage_times.py 2>&1 | sed -n 1,2p
