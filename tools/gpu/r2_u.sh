# round 2, call U: the whole GPU suite
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/u_tests.log; cat gpurun_out/u_tests.log
