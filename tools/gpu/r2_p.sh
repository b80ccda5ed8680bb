# round 2, call P: source-level profile of the block-per-instance kernel at batch 1
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ipm_solo -s 2 -c 1 -o gpurun_out/p_solo_b1 -f python tools/prof_k3.py 1 diff 4 > gpurun_out/p_ncu.log 2>&1; tail -2 gpurun_out/p_ncu.log
