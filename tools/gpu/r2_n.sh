# round 2, call N: hand-over threshold sweep after the omni4 factorising sweep got faster
mkdir -p gpurun_out
for f in 0.45 0.6 0.75 0.9; do
  for m in omni4 diff; do echo -n "frac $f: "; NMPC_HYB_FRAC=$f timeout 200 python tools/prof_k3.py 65536 $m 3 2>&1 | tail -1 | cut -c1-220; done
done | tee gpurun_out/n_frac.log
