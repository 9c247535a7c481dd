set -x
mkdir -p gpurun_out/r2b
timeout 90 python scripts/sanitize_case.py bf16x3 333 4096 64 1 20000 > gpurun_out/r2b/wide_first.log 2>&1; tail -8 gpurun_out/r2b/wide_first.log
timeout 90 python scripts/sanitize_case.py bf16 333 4096 > gpurun_out/r2b/wide_first_bf16.log 2>&1; tail -4 gpurun_out/r2b/wide_first_bf16.log
grep -q "sanitize_case: ok" gpurun_out/r2b/wide_first.log || exit 1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; tail -c 1500 gpurun_out/r2b/bench.json; tail -5 gpurun_out/r2b/bench.err
