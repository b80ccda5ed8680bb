# round 2, call O: first run of the block-per-instance kernel (rti_solo.cuh): batch-1 latency, small batches against the
# lane-cooperative kernel, parity tests
mkdir -p gpurun_out
timeout 120 python tools/prof_k3.py 1 diff 5 2>&1 | tail -1
timeout 120 python tools/prof_k3.py 1 tric 5 2>&1 | tail -1
for b in 1 16 148 296 512 1024 2048 4096; do
  for sm in 100000 0; do echo -n "solo_max=$sm: "; NMPC_SOLO_MAX=$sm timeout 120 python tools/prof_k3.py $b diff 4 2>&1 | tail -1 | cut -c1-230; done
done | tee gpurun_out/o_small.log
timeout 300 python tools/bench_models.py --latency-calls 500 --batches diff:1 2>&1 | grep latency | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_schedules.py tests/test_acados_dropin.py -x -q 2>&1 | tail -4
