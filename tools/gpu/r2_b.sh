# round 2, call B: occupancy / prefetch-depth variants of the lane-cooperative K3, ncu capture of the default build
mkdir -p gpurun_out
one() { echo -n "$1: "; env $2 NMPC_K3=group NMPC_GRP_IMPL=coop timeout 300 python tools/prof_k3.py 65536 ${3:-diff} 3 2>&1 | tail -1; }
one default "X=1"
one m2 "NMPC_B200_LIB=build/libnmpc_m2.so"
one m4 "NMPC_B200_LIB=build/libnmpc_m4.so"
one d3 "NMPC_B200_LIB=build/libnmpc_d3.so"
one omni4_default "X=1" omni4
one omni4_m2 "NMPC_B200_LIB=build/libnmpc_m2.so" omni4
one omni4_m4 "NMPC_B200_LIB=build/libnmpc_m4.so" omni4
echo -n "hybrid+coop: "; timeout 300 python tools/prof_k3.py 65536 diff 3 2>&1 | tail -1
echo -n "batch 4096: "; NMPC_K3=group timeout 300 python tools/prof_k3.py 4096 diff 3 2>&1 | tail -1
echo -n "batch 1: "; NMPC_K3=group timeout 300 python tools/prof_k3.py 1 diff 5 2>&1 | tail -1
NMPC_K3=group NMPC_GRP_IMPL=coop timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ipm_coop --launch-skip 1 --launch-count 1 -o gpurun_out/r2b_coop_full python tools/prof_k3.py 65536 diff 2 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
