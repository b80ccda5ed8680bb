# round 2, call W: config 4 through the block-per-instance kernel; phase cycles of omni4
timeout 600 python -m pytest tests/test_gpu_sqp_rollout.py -x -q -s -k config4 2>&1 | grep -E "config 4|passed|failed" | tail -3
NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py omni4 2>&1 | tail -26
