# N=<n> bash tools/gpu/run_mixed_multi.sh   (under gpurun --gpus <n>): BASELINE config 5 on n GPUs
mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_mixed_multi.py --total 1048576 > gpurun_out/mixed_n$N.json 2> gpurun_out/mixed_n$N.err
tail -1 gpurun_out/mixed_n$N.json | cut -c1-400; tail -2 gpurun_out/mixed_n$N.err | cut -c1-200
