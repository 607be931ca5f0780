set -x
timeout 600 python -m pytest tests/test_cli.py -m gpu -q -k "save" > gpurun_out/r2i_cli.log 2>&1; echo "cli rc=$?"; tail -20 gpurun_out/r2i_cli.log
