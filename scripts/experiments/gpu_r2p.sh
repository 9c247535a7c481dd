mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p/gpu_tests.log 2>&1; tail -4 gpurun_out/r2p/gpu_tests.log
timeout 300 python scripts/e2e_stress.py 100 64 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; tail -2 gpurun_out/r2p/bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2p/bench.json')); print('BENCH', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['int32_indices'].get('value'), d['parity']['max_rel'], d['roofline']['frac'], d['roofline']['stages'])"
