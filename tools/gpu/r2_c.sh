# round 2, call C: is the lane-cooperative K3 bound by its bulk throughput or by the straggler tail?
mkdir -p gpurun_out
one() { echo -n "$1: "; env $2 NMPC_K3=group NMPC_GRP_IMPL=coop timeout 300 python tools/prof_k3.py ${4:-65536} ${3:-diff} 3 2>&1 | tail -1; }
one default_e1 "X=1"
one e0 "NMPC_B200_LIB=build/libnmpc_e0.so"
one e0w2m7 "NMPC_B200_LIB=build/libnmpc_e0w2m7.so"
one e0w2m6 "NMPC_B200_LIB=build/libnmpc_e0w2m6.so"
one e0_itmax8 "NMPC_B200_LIB=build/libnmpc_e0.so PROF_ITER_MAX=8"
one e0_itmax5 "NMPC_B200_LIB=build/libnmpc_e0.so PROF_ITER_MAX=5"
one e0_131072 "NMPC_B200_LIB=build/libnmpc_e0.so" diff 131072
one e0_32768 "NMPC_B200_LIB=build/libnmpc_e0.so" diff 32768
one e0_16384 "NMPC_B200_LIB=build/libnmpc_e0.so" diff 16384
one e0_8192 "NMPC_B200_LIB=build/libnmpc_e0.so" diff 8192
for k in 2 3 4 5 6; do echo -n "hybrid kmax=$k frac=0: "; NMPC_B200_LIB=build/libnmpc_e0.so NMPC_HYB_KMAX=$k NMPC_HYB_FRAC=0.0 timeout 300 python tools/prof_k3.py 65536 diff 3 2>&1 | tail -1; done
