set -x
timeout 600 env MODES=megakernel python tools/ab.py libb200pt.so libb200pt_m7.so 2>&1 | tail -2
