mkdir -p gpurun_out/r2e
for s in 1 2 3 4 6 8; do timeout 200 python bench.py --steps 96 --warmup 5 --no-cpu-baseline --streams $s > gpurun_out/r2e/bench_s$s.json 2> gpurun_out/r2e/bench_s$s.err; python -c "
import json
d=json.load(open('gpurun_out/r2e/bench_s$s.json')); print('streams $s', round(d['value']/1e6,1), d['ms_per_step'], round(d['e2e']['value']/1e6,1), d['roofline']['ms_per_launch'])"; done
for B in 8192 16384 65536; do timeout 200 python bench.py --steps 48 --warmup 5 --no-cpu-baseline --batch $B --nbatches 32 > gpurun_out/r2e/bench_B$B.json 2> gpurun_out/r2e/bench_B$B.err; python -c "
import json
d=json.load(open('gpurun_out/r2e/bench_B$B.json')); print('B $B', round(d['value']/1e6,1), d['ms_per_step'], round(d['e2e']['value']/1e6,1), d['roofline']['ms_per_launch'])"; done
