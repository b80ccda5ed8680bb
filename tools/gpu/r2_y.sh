# round 2, call Y: the whole GPU suite + smoke
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/y_tests.log; cat gpurun_out/y_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
