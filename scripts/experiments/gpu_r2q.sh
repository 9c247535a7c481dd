mkdir -p gpurun_out/r2q
timeout 90 python scripts/sanitize_case.py bf16x3 333 4096 64 1 20000 > gpurun_out/r2q/wide.log 2>&1; tail -1 gpurun_out/r2q/wide.log
timeout 120 python scripts/wide_timeline.py 4096 bf16x3 > gpurun_out/r2q/tl4096.txt 2>&1; grep "back-to-back\|MMA: layer\|x_ready seen\|pair-tile 0\|shallow\|stored" gpurun_out/r2q/tl4096.txt | head -14
timeout 120 python scripts/wide_timeline.py 65536 bf16x3 > gpurun_out/r2q/tl65536.txt 2>&1; head -1 gpurun_out/r2q/tl65536.txt
for k in 20 100; do timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r2q/bench_$k.json 2> gpurun_out/r2q/bench_$k.err; python -c "
import json
d=json.load(open('gpurun_out/r2q/bench_$k.json')); print('BENCH K=$k', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1))"; done
