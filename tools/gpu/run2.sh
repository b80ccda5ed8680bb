mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t2.log
cat gpurun_out/t2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b_group2.log 2>&1
tail -1 gpurun_out/b_group2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['config']['mean_qp_iter'])"
