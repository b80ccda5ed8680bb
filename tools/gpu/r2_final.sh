# round 2, final evidence run (TAG names the files): bench line + reference arm, launch lists (time, DRAM bytes, executed fp64
# instructions), full ncu captures (sweeps, lane-cooperative kernel, block-per-instance kernel), secondary measurements
TAG=${TAG:-v2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err; tail -c 300 gpurun_out/r02_bench_$TAG.json; tail -3 gpurun_out/r02_bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_$TAG.json 2>> gpurun_out/r02_bench_$TAG.err; cut -c1-200 gpurun_out/r02_bench_reference_$TAG.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none -k regex:^k_ -c 100 --csv --log-file gpurun_out/r02_launches_hybrid_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sweep --launch-skip 3 --launch-count 2 -o gpurun_out/r02_sweeps_full_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_sw.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ipm_coop -c 1 -o gpurun_out/r02_coop_resume_full_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_coop.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ipm_solo -s 2 -c 1 -o gpurun_out/r02_solo_b1_full_$TAG -f python tools/prof_k3.py 1 diff 4 > gpurun_out/ncu_solo.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:^k_ -c 45 --csv --log-file gpurun_out/r02_launches_omni4_$TAG.csv python tools/bench_models.py --latency-calls 0 --batches omni4:65536 > gpurun_out/ncu_omni.log 2>&1
timeout 1200 python tools/bench_models.py --latency-calls 1000 --batches diff:1,diff:148,diff:592,diff:1024,diff:4096,diff:16384,diff:65536,diff:131072,tric:148,tric:65536,omni4:65536 > gpurun_out/r02_models_$TAG.jsonl 2> gpurun_out/r02_models_$TAG.err; tail -4 gpurun_out/r02_models_$TAG.jsonl | cut -c1-200
timeout 600 python tools/bench_models.py --ctrl diff:65536,tric:65536,omni4:65536,diff:1 > gpurun_out/r02_ctrl_$TAG.jsonl 2> gpurun_out/r02_ctrl_$TAG.err; cut -c1-200 gpurun_out/r02_ctrl_$TAG.jsonl
[ -f build/var/lib_soloprof.so ] && NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py diff > gpurun_out/r02_solo_phase_cycles_$TAG.txt 2>&1
ls -la gpurun_out | tail -12
