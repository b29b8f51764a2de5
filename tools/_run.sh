set -x
timeout 300 python tools/stage_times.py 2>&1 | sed -n 1,2p
timeout 300 env PT_LIB_PATH=/root/repo/path_tracer_python_b200/libb200pt_s7.so python tools/stage_times.py 2>&1 | sed -n 1,2p
timeout 300 python tools/stage_times.py 2>&1 | sed -n 1,2p
timeout 300 env PT_LIB_PATH=/root/repo/path_tracer_python_b200/libb200pt_s7.so python tools/stage_times.py 2>&1 | sed -n 1,2p
