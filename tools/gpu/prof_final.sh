mkdir -p gpurun_out
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 600 gpurun_out/bench_r01_final.json
# launch list of the bench command (per-launch time and DRAM bytes; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:^k_ -c 120 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
# full captures: FDF and B sweeps of the first iteration, the group kernel
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sweep --launch-skip 1 --launch-count 2 -o gpurun_out/sweeps_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_sw.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ipm_group -c 1 -o gpurun_out/grp_resume_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_grp.log 2>&1
ls -la gpurun_out
timeout 1200 python tools/bench_models.py --latency-calls 1000 > gpurun_out/models_final.jsonl 2> gpurun_out/models_final.err
timeout 1200 python tools/bench_models.py --mixed 1048576 >> gpurun_out/models_final.jsonl 2>> gpurun_out/models_final.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:^k_ -c 45 --csv --log-file gpurun_out/launches_omni4.csv python tools/bench_models.py --latency-calls 1 --batches omni4:65536 > gpurun_out/ncu_omni.log 2>&1
# controller tick and path discretisation
timeout 600 python tools/bench_models.py --ctrl diff:65536,tric:65536,omni4:65536,diff:1 > gpurun_out/ctrl_final.jsonl 2> gpurun_out/ctrl_final.err
timeout 300 python tools/bench_models.py --paths 65536 >> gpurun_out/ctrl_final.jsonl 2>> gpurun_out/ctrl_final.err
timeout 300 python tools/bench_models.py --paths 1048576 >> gpurun_out/ctrl_final.jsonl 2>> gpurun_out/ctrl_final.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_ctrl|k_path" -c 12 --csv --log-file gpurun_out/launches_ctrl.csv python tools/bench_models.py --ctrl diff:65536 > gpurun_out/ncu_ctrl.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_path" -c 6 --csv --log-file gpurun_out/launches_path.csv python tools/bench_models.py --paths 65536 > gpurun_out/ncu_path.log 2>&1
