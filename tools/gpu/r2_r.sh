set -x
timeout 600 python -m pytest tests/test_gpu_hostpipe.py -m gpu -q -x > gpurun_out/r2r_pipe.log 2>&1; tail -5 gpurun_out/r2r_pipe.log
timeout 900 python bench.py > gpurun_out/r2r_bench1.json 2> gpurun_out/r2r_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2r_bench1.err
python - <<'PY'
import json
for line in open('gpurun_out/r2r_bench1.json'):
    if line.startswith('{'):
        d=json.loads(line); print('%.4g'%d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:1500]); print(json.dumps(d['roofline'])[:800]); print(json.dumps(d.get('extra'))[:1500]); print(json.dumps(d.get('cpu_baseline'))[:600])
PY
