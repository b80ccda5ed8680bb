mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_sqp_rollout.py tests/test_gpu_parity.py::test_benchmarked_size_and_schedule_against_the_oracle tests/test_gpu_emit.py -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2f_tests.log; grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2f_tests.log | tail -15
grep -E "(SQP|rollout engine|config 4|closed loop|diff B=|tric B=|omni4 B=)" gpurun_out/r2f_tests.log | head -40
