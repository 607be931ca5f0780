set -x
timeout 900 python -m pytest tests/test_gpu_penalty.py tests/test_gpu_mono.py tests/test_gpu_fa.py tests/test_gpu_stage.py -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2l_tests.log
