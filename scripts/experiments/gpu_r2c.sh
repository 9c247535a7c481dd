set -x
mkdir -p gpurun_out/r2c
timeout 90 python scripts/sanitize_case.py bf16x3 333 4096 64 1 20000 > gpurun_out/r2c/wide.log 2>&1; tail -3 gpurun_out/r2c/wide.log
grep -q "sanitize_case: ok" gpurun_out/r2c/wide.log || exit 1
timeout 120 python scripts/wide_timeline.py 4096 bf16x3 > gpurun_out/r2c/timeline_4096.txt 2>&1; head -32 gpurun_out/r2c/timeline_4096.txt
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r2c/timeline_65536.txt 2>&1; sed -n 1,3p gpurun_out/r2c/timeline_65536.txt; sed -n 29,58p gpurun_out/r2c/timeline_65536.txt
timeout 300 python bench.py --steps 96 --warmup 5 --no-cpu-baseline --streams 4 > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2c/bench.json')); print('BENCH', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['ms_per_launch'])"
