mkdir -p gpurun_out/r2w
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/r2w/pytest.log 2>&1; tail -2 gpurun_out/r2w/pytest.log
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r2w/tl65536.txt 2>&1; head -1 gpurun_out/r2w/tl65536.txt; sed -n '/tile 1/,/tile 2/p' gpurun_out/r2w/tl65536.txt | grep "gather\|MMA"
for k in 20 100; do timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r2w/bench_$k.json 2> gpurun_out/r2w/bench_$k.err; python -c "
import json
d=json.load(open('gpurun_out/r2w/bench_$k.json')); print('BENCH K=$k', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['clocks'])"; done
timeout 300 python bench.py --workload twitter --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2w/bench_tw.json 2> gpurun_out/r2w/bench_tw.err; python -c "
import json
d=json.load(open('gpurun_out/r2w/bench_tw.json')); print('BENCH tw', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1))"
