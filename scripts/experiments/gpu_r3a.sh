mkdir -p gpurun_out/r3a
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r3a/tl65536.txt 2>&1; head -1 gpurun_out/r3a/tl65536.txt; grep "x_ready seen\|x_ready arr\|X released\|layer 1 issued\|layer 3 issued" gpurun_out/r3a/tl65536.txt | sed -n 1,24p
timeout 500 python -m pytest tests/ -m gpu -x -q > gpurun_out/r3a/pytest.log 2>&1; tail -2 gpurun_out/r3a/pytest.log
timeout 300 python scripts/e2e_stress.py 150 64 > gpurun_out/r3a/stress.log 2>&1; tail -1 gpurun_out/r3a/stress.log
for k in 20 100; do timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r3a/bench_$k.json 2> gpurun_out/r3a/bench_$k.err; python -c "
import json
d=json.load(open('gpurun_out/r3a/bench_$k.json')); print('BENCH K=$k', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['parity']['max_rel'], d['roofline']['frac'], {k:v.get('samples_per_s') for k,v in d['roofline']['stages'].items() if '65536' in k})"; done
