mkdir -p gpurun_out
set -x
B=${B:-65536}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ipm_group -c 1 -o gpurun_out/grp_full python bench.py --steps 1 --warmup 3 --no-cpu --batch $B > gpurun_out/ncu_grp.log 2>&1
tail -3 gpurun_out/ncu_grp.log | cut -c1-300
ls -la gpurun_out
