set -x
mkdir -p gpurun_out/r2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_plain.json 2> gpurun_out/r2/bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/launches_bf16x3.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 6 -c 2 -o gpurun_out/r2/full_bf16x3 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_full.log 2>&1
ls -la gpurun_out/r2/
