# round 2, call H: the bench contract under torch.distributed.run (2 ranks): both arms
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 1500 gpurun_out/r02_bench_2gpu.json; tail -3 gpurun_out/r02_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_2gpu.json 2> gpurun_out/r02_bench_reference_2gpu.err; cut -c1-600 gpurun_out/r02_bench_reference_2gpu.json; tail -3 gpurun_out/r02_bench_reference_2gpu.err
