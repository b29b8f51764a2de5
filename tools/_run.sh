set -x
timeout 400 env SPP=64 SNODES=-1 BURSTS=8,12,16,20,24 VOTES=8,12,16,20 python tools/sweep.py 2>&1 | tail -21
timeout 300 env BURSTS=16,20,24 VOTES=12,16,20 python tools/sweep_wf.py 2>&1 | tail -9
timeout 300 env SNODES=64,128,256,512 python tools/sweep_wf.py 2>&1 | tail -4
