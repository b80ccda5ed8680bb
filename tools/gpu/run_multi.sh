mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
tail -2 gpurun_out/bench_n$N.err | cut -c1-200
