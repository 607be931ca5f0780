timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/r2_final_bench1.json 2> gpurun_out/r2_final_bench1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_final_bench1.err
python - <<'PY'
import json
for line in open('gpurun_out/r2_final_bench1.json'):
    if line.startswith('{'):
        d=json.loads(line); print('%.4g'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks']); print(json.dumps(d['e2e'])[:700]); print(json.dumps(d.get('extra'))[:1200]); print(json.dumps(d.get('cpu_baseline'))[:400])
PY
