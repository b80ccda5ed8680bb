mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t6.log; cat gpurun_out/t6.log
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'qp_ms', round(d['kernel_ms']['qp_ms'],2), 'launches', d['gpu_launches'], 'it', d['config']['mean_qp_iter'], 'bad', d['config']['status_nonzero'])"; }
run hybrid_default "X=1"
run group "NMPC_K3=group"
timeout 900 python tools/bench_models.py --latency-calls 300 --batches diff:4096,diff:16384,tric:65536,omni4:65536 > gpurun_out/models.jsonl 2> gpurun_out/models.err; tail -3 gpurun_out/models.err
python - <<'PY'
import json
for l in open('gpurun_out/models.jsonl'):
    d=json.loads(l)
    if d['kind']=='throughput': print(d['model'],d['batch'],round(d['ms_per_step'],2),'ms', round(d['solves_per_s']), 'solves/s', 'it',round(d['mean_qp_iter'],2),d['max_qp_iter'],'bad',d['status_nonzero'], {k:round(v,2) for k,v in d['kernel_ms'].items()})
    else: print(d['model'],'batch-1 latency us p50/p95/p99',round(d['p50_us']),round(d['p95_us']),round(d['p99_us']),'qp_iter',d['qp_iter'])
PY
