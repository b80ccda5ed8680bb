# round 2, call I: baseline of the session + batch-1 source-level profile of the lane-cooperative kernel + group-only schedules
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; cut -c1-300 gpurun_out/i_bench.json
for m in diff omni4 tric; do NMPC_K3=group timeout 200 python tools/prof_k3.py 65536 $m 3 2>&1 | tail -1; done > gpurun_out/i_group_only.log; cat gpurun_out/i_group_only.log
timeout 100 python tools/prof_k3.py 1 diff 5 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ipm_coop -s 3 -c 1 -o gpurun_out/i_coop_b1 -f python tools/prof_k3.py 1 diff 5 > gpurun_out/i_ncu_b1.log 2>&1; tail -2 gpurun_out/i_ncu_b1.log
