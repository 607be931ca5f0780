set -x
timeout 900 python -m pytest tests/test_gpu_penalty.py -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2h_tests.log
timeout 600 python -m pytest tests/test_cli.py -m gpu -q -k "penalty or mono" > gpurun_out/r2h_cli.log 2>&1; echo "cli rc=$?"; tail -20 gpurun_out/r2h_cli.log
