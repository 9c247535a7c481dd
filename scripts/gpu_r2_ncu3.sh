set -x
mkdir -p gpurun_out/r2f
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/bench_plain.json 2> gpurun_out/r2f/bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f/launches_bf16x3.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 6 -c 2 -o gpurun_out/r2f/full_bf16x3 -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 2 -c 1 -o gpurun_out/r2f/full_b65536_bf16x3 -f python scripts/run_big_batch.py 65536 bf16x3 4 > gpurun_out/r2f/ncu_full_b65536.log 2>&1
ls -la gpurun_out/r2f/
