timeout 900 python -m pytest tests -m gpu -q -x -k "remap or product or cli or golden or known" > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ae_tests.log
python - <<'PY'
import json, sys
sys.argv=['bench.py']
import bench
r = bench.extra_runs(3, 5, 0)
print(json.dumps(r['remap']))
PY
