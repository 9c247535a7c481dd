mkdir -p gpurun_out/r2k
for w in criteo criteo_qr_small; do
timeout 300 ncu --set full --clock-control none -k regex:fused_wide -s 2 -c 1 -o gpurun_out/r2k/full_$w -f python scripts/run_one.py $w > gpurun_out/r2k/runfull_$w.log 2>&1
done
ls gpurun_out/r2k
