# round 2: block-per-instance kernel: latency of the default build, phase cycles of the -DNMPC_SOLO_PROF build
timeout 120 python tools/prof_k3.py 1 diff 5 2>&1 | tail -1 | cut -c1-250
timeout 120 python tools/prof_k3.py 1 tric 5 2>&1 | tail -1 | cut -c1-250
NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py diff 2>&1 | tail -26
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
