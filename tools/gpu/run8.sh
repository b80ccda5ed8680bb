for f in 0.2 0.3 0.4 0.5 0.6; do
  echo -n "FRAC=$f: "
  NMPC_HYB_FRAC=$f timeout 900 python tools/bench_models.py --latency-calls 1 --batches diff:65536,tric:65536,omni4:65536 2>/dev/null | python -c "
import sys,json
o=[]
for l in sys.stdin:
    d=json.loads(l)
    if d['kind']=='throughput': o.append('%s %.2f ms'%(d['model'],d['ms_per_step']))
print(' | '.join(o))
"
done
