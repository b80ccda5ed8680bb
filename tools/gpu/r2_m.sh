# round 2, call M: lean view of [A B] in the omni4 Riccati half; parity tests of the default build
mkdir -p gpurun_out
for m in omni4 diff tric; do timeout 200 python tools/prof_k3.py 65536 $m 4 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
