set -x
for l in libb200pt.so libb200pt_v3.so; do
timeout 300 env PT_LIB_PATH=/root/repo/path_tracer_python_b200/$l python tools/stage_times.py 2>&1 | tail -5
done
