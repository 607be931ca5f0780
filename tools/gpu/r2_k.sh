set -x
timeout 900 python -m pytest tests/test_cli.py -m gpu -q -k "decomposed" > gpurun_out/r2k_cli.log 2>&1; echo "cli rc=$?"; tail -40 gpurun_out/r2k_cli.log
