mkdir -p gpurun_out/r2m
timeout 600 python scripts/stress_repro.py 500 > gpurun_out/r2m/stress_repro.log 2>&1; cat gpurun_out/r2m/stress_repro.log | tail -14
