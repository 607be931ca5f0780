set -x
python bench.py --nloc 96 --steps 100 --no-extras --no-cpu-baseline > gpurun_out/r2n_cart.json 2>/dev/null
python bench.py --steps 100 --no-extras --no-cpu-baseline > gpurun_out/r2n_ref.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r2n_cart.json','gpurun_out/r2n_ref.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])
PY
