mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
