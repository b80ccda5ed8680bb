# hybrid hand-over threshold sweep per model (NMPC_HYB_FRAC, NMPC_HYB_KMAX); prints ms per step
mkdir -p gpurun_out
for f in ${FRACS:-0.1 0.2 0.3 0.4 0.55}; do
  NMPC_HYB_FRAC=$f timeout 300 python tools/bench_models.py --latency-calls 0 --batches diff:65536,tric:65536,omni4:65536 2>/dev/null |
    python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('frac $f', d['model'], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})"
done | tee gpurun_out/sweep_frac.txt
