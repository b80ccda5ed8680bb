# round 2, call Q: block-per-instance kernel after the block-wide Riccati: latency, phase cycles, parity
mkdir -p gpurun_out
timeout 120 python tools/prof_k3.py 1 diff 5 2>&1 | tail -1 | cut -c1-250
timeout 120 python tools/prof_k3.py 1 tric 5 2>&1 | tail -1 | cut -c1-250
NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py diff 2>&1 | tail -18
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
