mkdir -p gpurun_out/r2z
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r2z/tl65536.txt 2>&1; head -1 gpurun_out/r2z/tl65536.txt; grep "x_ready seen\|pair-tile" gpurun_out/r2z/tl65536.txt | sed -n 14,30p
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2z/pytest.log 2>&1; tail -1 gpurun_out/r2z/pytest.log
for k in 100; do timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r2z/bench_$k.json 2> gpurun_out/r2z/bench_$k.err; python -c "
import json
d=json.load(open('gpurun_out/r2z/bench_$k.json')); print('BENCH K=$k', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['parity']['max_rel'], d['roofline']['frac'])"; done
