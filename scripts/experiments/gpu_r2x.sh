mkdir -p gpurun_out/r2x
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r2x/tl65536.txt 2>&1; head -1 gpurun_out/r2x/tl65536.txt; grep "XEPI\|x_ready seen\|X released\|x_ready arr" gpurun_out/r2x/tl65536.txt
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/r2x/pytest.log 2>&1; tail -2 gpurun_out/r2x/pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2x/bench_100.json 2> gpurun_out/r2x/bench_100.err; python -c "
import json
d=json.load(open('gpurun_out/r2x/bench_100.json')); print('BENCH K=100', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['parity']['max_rel'])"
