set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2d_tests.log
