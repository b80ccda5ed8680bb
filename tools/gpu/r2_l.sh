# round 2, call L: launch list (times + DRAM bytes) of an omni4 step with the streamed update half
mkdir -p gpurun_out
timeout 200 python tools/prof_k3.py 65536 omni4 3 2>&1 | tail -1
timeout 200 python tools/prof_k3.py 65536 diff 3 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_sweep|k_ipm|k_handover|k_linearize|k_step" -c 40 --csv --log-file gpurun_out/l_launch_omni4.csv python tools/prof_k3.py 65536 omni4 1 > /dev/null 2>&1
python tools/ncu_launches.py gpurun_out/l_launch_omni4.csv 2>&1 | head -45
