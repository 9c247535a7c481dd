mkdir -p gpurun_out/r2s
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "qr or QR" > gpurun_out/r2s/pytest_qr.log 2>&1; tail -2 gpurun_out/r2s/pytest_qr.log
timeout 120 python scripts/wide_timeline.py 4096 bf16x3 criteo_qr > gpurun_out/r2s/tl_qr4096.txt 2>&1; grep "back-to-back\|rows of both\|x_ready arrive\|shallow\|stored" gpurun_out/r2s/tl_qr4096.txt | head -14
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 criteo_qr > gpurun_out/r2s/tl_qr65536.txt 2>&1; head -1 gpurun_out/r2s/tl_qr65536.txt; grep "rows of both" gpurun_out/r2s/tl_qr65536.txt
timeout 300 python bench.py --workload criteo_qr --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2s/bench_qr.json 2> gpurun_out/r2s/bench_qr.err; python -c "
import json
d=json.load(open('gpurun_out/r2s/bench_qr.json')); print('BENCH qr', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d.get('parity'))"
