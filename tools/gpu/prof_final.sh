mkdir -p gpurun_out
set -x
# launch list of the bench command (per-launch time and DRAM bytes; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 400 --csv --log-file gpurun_out/launches_group.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
# full capture of the K3 kernel
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ipm_group -c 1 -o gpurun_out/grp_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_grp.log 2>&1
ls -la gpurun_out
