mkdir -p gpurun_out/r2
python scripts/run_big_batch.py 65536 bf16x3 4 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 2 -c 1 -o gpurun_out/r2/full_b65536_bf16x3 -f python scripts/run_big_batch.py 65536 bf16x3 4 > gpurun_out/r2/ncu_full_b65536.log 2>&1
ls -la gpurun_out/r2/
