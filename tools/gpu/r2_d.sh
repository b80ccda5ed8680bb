# round 2, call D: the new GPU tests (SQP / shift / rollout engine / full-size parity), the hybrid default, the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sqp_rollout.py tests/test_gpu_parity.py tests/test_acados_dropin.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2d_tests.log; cat gpurun_out/r2d_tests.log
echo -n "hybrid default: "; timeout 300 python tools/prof_k3.py 65536 diff 3 2>&1 | tail -1
echo -n "hybrid default tric: "; timeout 300 python tools/prof_k3.py 65536 tric 3 2>&1 | tail -1
echo -n "hybrid default omni4: "; timeout 300 python tools/prof_k3.py 65536 omni4 3 2>&1 | tail -1
echo -n "coop omni4: "; NMPC_K3=group timeout 300 python tools/prof_k3.py 65536 omni4 3 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 3000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
