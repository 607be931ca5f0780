set -x
timeout 900 python -m pytest tests/test_gpu_product.py tests/test_gpu_scale.py -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/r2e_tests.log
timeout 600 python -m pytest tests/test_cli.py -m gpu -q -k "product or default" > gpurun_out/r2e_cli.log 2>&1; echo "cli rc=$?"; tail -30 gpurun_out/r2e_cli.log
