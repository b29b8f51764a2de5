set -x
timeout 900 python tools/ab.py libb200pt.so libb200pt_gen.so 2>&1 | tail -5
timeout 300 env SNODES=0,256,1024,2048 python tools/sweep_wf.py 2>&1 | tail -5
