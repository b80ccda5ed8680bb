# round 2, call E: the whole GPU suite with its printed statistics
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2e_tests_full.log; grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2e_tests_full.log | tail -15
grep -E "^(SQP|rollout engine|config 4|closed loop|diff B=|tric B=|omni4 B=|diff:|tric:|omni4:)" gpurun_out/r2e_tests_full.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
