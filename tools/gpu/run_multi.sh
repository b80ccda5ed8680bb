mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
tail -3 gpurun_out/bench_n$N.err
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
