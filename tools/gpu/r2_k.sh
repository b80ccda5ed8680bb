# round 2, call K: baseline library (HEAD) against the channel-streamed update (v0) with the same tool; kernel times by ncu
mkdir -p gpurun_out
for v in base v0; do
  for m in diff tric; do
    echo -n "$v $m: "; NMPC_B200_LIB=$PWD/build/var/lib_$v.so timeout 200 python tools/prof_k3.py 65536 $m 4 2>&1 | tail -1
  done
done > gpurun_out/k_variants.log 2>&1
cat gpurun_out/k_variants.log
for v in base v0; do
  NMPC_B200_LIB=$PWD/build/var/lib_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/k_launch_$v.csv python tools/prof_k3.py 65536 diff 1 > /dev/null 2>&1
  python - gpurun_out/k_launch_$v.csv <<'P'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[ki][:40]; v=float(r[vi].replace(',',''))/1e6
    agg.setdefault(n,[]).append(v)
for n,v in agg.items():
    if sum(v)>0.05: print(sys.argv[1][-12:], n, len(v), 'sum %.3f ms'%sum(v), 'each', ' '.join('%.2f'%x for x in v[:8]))
P
done
