mkdir -p gpurun_out/tests
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests/gpu_tests.log 2>&1; tail -15 gpurun_out/tests/gpu_tests.log
