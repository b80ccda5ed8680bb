mkdir -p gpurun_out
timeout 900 python tools/bench_models.py > gpurun_out/models.jsonl 2> gpurun_out/models.err; tail -3 gpurun_out/models.err
python - <<'PY'
import json
for l in open('gpurun_out/models.jsonl'):
    d=json.loads(l)
    if d['kind']=='throughput': print(d['model'],d['batch'],round(d['ms_per_step'],2),'ms', round(d['solves_per_s']), 'solves/s', 'it',round(d['mean_qp_iter'],2),d['max_qp_iter'],'bad',d['status_nonzero'], {k:round(v,2) for k,v in d['kernel_ms'].items()})
    else: print(d)
PY
