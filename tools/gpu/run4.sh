mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t4.log; cat gpurun_out/t4.log
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'qp_ms', round(d['kernel_ms']['qp_ms'],2), 'launches', d['gpu_launches'], 'it', d['config']['mean_qp_iter'], 'bad', d['config']['status_nonzero'])"; }
run hybrid_default "X=1"
for v in $VARIANTS; do run $v "$v"; done
