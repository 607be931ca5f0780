set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_full_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_full_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
