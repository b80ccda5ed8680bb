mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/bench_models.py --latency-calls 1 --batches diff:65536,tric:65536,omni4:65536,omni4:262144 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['kind']=='throughput': print(' ',d['model'],d['batch'],round(d['ms_per_step'],2),'ms',round(d['solves_per_s']),'solves/s')
"
