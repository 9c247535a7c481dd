set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L
python scripts/precision_survey.py > gpurun_out/r2a/precision_survey.txt 2>&1
python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/r2a/fullsize_tests.log 2>&1; tail -5 gpurun_out/r2a/fullsize_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a/bench_criteo.json 2> gpurun_out/r2a/bench_criteo.err; tail -c 3000 gpurun_out/r2a/bench_criteo.json
for p in bf16x3 fp32_csr bf16 fp32; do python bench.py --workload criteo_pruned --precision $p --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_pruned_$p.json 2> gpurun_out/r2a/bench_pruned_$p.err; done
tail -c 600 gpurun_out/r2a/bench_pruned_*.json
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py bf16x3 4096 333 > gpurun_out/r2a/sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/r2a/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_case.py bf16x3 4096 333 > gpurun_out/r2a/sanitizer_racecheck.log 2>&1; tail -5 gpurun_out/r2a/sanitizer_racecheck.log
