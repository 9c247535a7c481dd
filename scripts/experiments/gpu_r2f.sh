mkdir -p gpurun_out/r2f
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "host or streamed or transports or loader or tiny or odd_widths" > gpurun_out/r2f/tests.log 2>&1; tail -5 gpurun_out/r2f/tests.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; tail -3 gpurun_out/r2f/bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2f/bench.json')); print('BENCH', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['int32_indices'], d['roofline']['stages'])"
