set -x
timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
python bench.py --steps 200 --no-extras --no-cpu-baseline > gpurun_out/r2p_bench1.json 2>/dev/null
python bench.py --steps 200 --order 4 --no-extras --no-cpu-baseline > gpurun_out/r2p_bench1_o4.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2p_bench1.json','gpurun_out/r2p_bench1_o4.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['check'])
PY
