# round 2, call T: the whole GPU suite with the block-per-instance kernel in place; small-batch crossover; latency
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_tests.log; cat gpurun_out/t_tests.log
for b in 1 148 296 444 592 740 1024; do
  for sm in 100000 0; do echo -n "solo_max=$sm: "; NMPC_SOLO_MAX=$sm timeout 120 python tools/prof_k3.py $b diff 4 2>&1 | tail -1 | sed -E "s/status.*'qp_ms'/qp_ms/" | cut -c1-150; done
done | tee gpurun_out/t_small.log
timeout 300 python tools/bench_models.py --latency-calls 1000 --batches diff:1 2>&1 | grep latency | cut -c1-260 | tee gpurun_out/t_latency.jsonl
