mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print('bench default', round(d['value']), d['steps'], d['warmup'], round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'clocks', d['clocks'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
