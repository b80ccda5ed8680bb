mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t5.log; cat gpurun_out/t5.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python tools/bench_models.py --batches diff:8192,diff:16384,diff:32768,diff:65536,diff:131072,tric:65536,omni4:65536 > gpurun_out/models.jsonl 2> gpurun_out/models.err; tail -3 gpurun_out/models.err
NMPC_K3=group timeout 900 python tools/bench_models.py --latency-calls 10 --batches diff:8192,diff:16384,diff:32768,omni4:65536 > gpurun_out/models_group.jsonl 2>> gpurun_out/models.err
NMPC_HYB_MIN=0 timeout 900 python tools/bench_models.py --latency-calls 10 --batches diff:8192,diff:16384,diff:32768 > gpurun_out/models_hyb.jsonl 2>> gpurun_out/models.err
python - <<'PY'
import json
for f in ('models','models_group','models_hyb'):
    print('==',f)
    for l in open('gpurun_out/%s.jsonl'%f):
        d=json.loads(l)
        if d['kind']=='throughput': print(d['model'],d['batch'],round(d['ms_per_step'],2),'ms', round(d['solves_per_s']), 'solves/s', 'it',round(d['mean_qp_iter'],2),d['max_qp_iter'],'bad',d['status_nonzero'], {k:round(v,2) for k,v in d['kernel_ms'].items()})
        elif f=='models': print(d['model'],'batch-1 latency us p50/p95/p99',round(d['p50_us']),round(d['p95_us']),round(d['p99_us']),'qp_iter',d['qp_iter'])
PY
