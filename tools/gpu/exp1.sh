mkdir -p gpurun_out
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'qp_ms', round(d['kernel_ms']['qp_ms'],2))"; }
run base "X=1"
run pad0 "NMPC_B200_LIB=$PWD/tools/gpu/exp/lib_pad0.so"
run pad4 "NMPC_B200_LIB=$PWD/tools/gpu/exp/lib_pad4.so"
run pad12 "NMPC_B200_LIB=$PWD/tools/gpu/exp/lib_pad12.so"
run bps2 "NMPC_GRP_BPS=2"
run bps1 "NMPC_GRP_BPS=1"
run G16 "NMPC_GRP_G=16"
run G32 "NMPC_GRP_G=32"
