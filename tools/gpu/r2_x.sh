# round 2, call X: block-per-instance kernel after the row / [B A] prefetch: the three models at batch 1, omni4 phase cycles, tests
for m in diff tric omni4; do timeout 120 python tools/prof_k3.py 1 $m 5 2>&1 | tail -1 | sed -E "s/status.*'qp_ms'/qp_ms/" | cut -c1-150; done
for b in 148 296; do for sm in 100000 0; do echo -n "solo_max=$sm: "; NMPC_SOLO_MAX=$sm timeout 120 python tools/prof_k3.py $b omni4 4 2>&1 | tail -1 | sed -E "s/status.*'qp_ms'/qp_ms/" | cut -c1-150; done; done
NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py omni4 2>&1 | tail -26
timeout 900 python -m pytest tests/test_gpu_solo.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
