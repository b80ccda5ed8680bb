mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t1.log
cat gpurun_out/t1.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b_group.log 2>&1
tail -3 gpurun_out/b_group.log
NMPC_K3=sweep timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b_sweep.log 2>&1
tail -1 gpurun_out/b_sweep.log | cut -c1-400
