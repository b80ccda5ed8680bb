# round 2, call V: omni4 through the block-per-instance kernel (global scratch for [B A], Phi), SQP passes through it, tests
mkdir -p gpurun_out
for m in omni4; do for b in 1 16 148 296; do for sm in 100000 0; do echo -n "solo_max=$sm: "; NMPC_SOLO_MAX=$sm timeout 120 python tools/prof_k3.py $b $m 4 2>&1 | tail -1 | sed -E "s/status.*'qp_ms'/qp_ms/" | cut -c1-150; done; done; done | tee gpurun_out/v_omni4.log
timeout 1500 python -m pytest tests/test_gpu_solo.py tests/test_gpu_parity.py tests/test_gpu_sqp_rollout.py tests/test_gpu_controller.py tests/test_gpu_rollout.py tests/test_acados_dropin.py -x -q -s 2>&1 | grep -E "config 4|passed|failed|Error|error" | tail -8
NMPC_SOLO_MAX=0 timeout 600 python -m pytest tests/test_gpu_sqp_rollout.py -x -q -s -k config4 2>&1 | grep -E "config 4|passed|failed" | tail -3
