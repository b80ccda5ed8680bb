# round 2, call A: first run of the lane-cooperative K3 (rti_coop.cuh) on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_schedules.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_tests.log; cat gpurun_out/r2a_tests.log
run() { echo "== $1"; env $2 timeout 600 python tools/bench_models.py --latency-calls $3 --batches "$4" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.rstrip()[:200]); continue
    if d['kind'] == 'throughput': print(d['model'], d['batch'], round(d['ms_per_step'], 2), 'ms', round(d['solves_per_s']), 'solves/s', 'it', round(d['mean_qp_iter'], 2), d['max_qp_iter'], 'bad', d['status_nonzero'], {k: round(v, 2) for k, v in d['kernel_ms'].items()})
    else: print(d['model'], 'batch-1 latency us p50/p95/p99', round(d['p50_us']), round(d['p95_us']), round(d['p99_us']), 'qp_iter', d['qp_iter'])
"; }
run "coop alone"   "NMPC_K3=group NMPC_GRP_IMPL=coop"  200 "diff:4096,diff:65536,tric:65536,omni4:65536"
run "group alone"  "NMPC_K3=group NMPC_GRP_IMPL=group" 200 "diff:4096,diff:65536,omni4:65536"
run "hybrid+coop"  "NMPC_GRP_IMPL=coop"  0 "diff:65536,tric:65536,omni4:65536"
run "hybrid+group" "NMPC_GRP_IMPL=group" 0 "diff:65536"
