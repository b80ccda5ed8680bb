# round 2, call G: bench line, launch list (time, DRAM bytes, executed fp64 instructions), full captures, secondary measurements
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 400 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench_reference.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none -k regex:^k_ -c 100 --csv --log-file gpurun_out/r02_launches_hybrid.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sweep --launch-skip 3 --launch-count 2 -o gpurun_out/r02_sweeps_full python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_sw.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ipm_coop -c 1 -o gpurun_out/r02_coop_resume_full python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_coop.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:^k_ -c 45 --csv --log-file gpurun_out/r02_launches_omni4.csv python tools/bench_models.py --latency-calls 0 --batches omni4:65536 > gpurun_out/ncu_omni.log 2>&1
timeout 1200 python tools/bench_models.py --latency-calls 1000 --batches diff:1024,diff:4096,diff:16384,diff:65536,diff:131072,tric:65536,omni4:65536 > gpurun_out/r02_models.jsonl 2> gpurun_out/r02_models.err; tail -4 gpurun_out/r02_models.jsonl | cut -c1-200
timeout 600 python tools/bench_models.py --ctrl diff:65536,tric:65536,omni4:65536,diff:1 > gpurun_out/r02_ctrl.jsonl 2> gpurun_out/r02_ctrl.err; cut -c1-200 gpurun_out/r02_ctrl.jsonl
ls -la gpurun_out | tail -20
