mkdir -p gpurun_out/full
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full/gpu_tests.log 2>&1; tail -4 gpurun_out/full/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full/smoke.log 2>&1; tail -4 gpurun_out/full/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/full/bench_reference.json 2> gpurun_out/full/bench_reference.err; cut -c1-300 gpurun_out/full/bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/full/bench.json 2> gpurun_out/full/bench.err; tail -3 gpurun_out/full/bench.err; python -c "
import json
d=json.load(open('gpurun_out/full/bench.json')); print('BENCH', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['int32_indices'].get('value'), d['parity']['max_rel'], d['cpu_baseline']['value'], d['reference_cuda'].get('value'), d['clocks'])"
