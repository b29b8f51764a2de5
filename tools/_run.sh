set -x
timeout 600 python tools/exp_sort.py 2>&1 | tail -20
