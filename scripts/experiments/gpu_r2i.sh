mkdir -p gpurun_out/r2i
for s in 4 5 6 8 10; do timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --streams $s > gpurun_out/r2i/bench_s$s.json 2> gpurun_out/r2i/bench_s$s.err; python -c "
import json
d=json.load(open('gpurun_out/r2i/bench_s$s.json')); print('K=20 streams $s', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['e2e']['int32_indices'].get('value'))"; done
