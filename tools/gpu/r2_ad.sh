timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 200 > gpurun_out/r2ad_bench4.json 2> gpurun_out/r2ad_bench4.err; echo "bench4 rc=$?"; tail -2 gpurun_out/r2ad_bench4.err
python - <<'PY'
import json
for f in ('gpurun_out/r2ad_bench4.json',):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'].get('dist_rel_err'), d['check']['mass_rel_drift'], json.dumps(d['e2e'])[:900]); print(json.dumps(d['halo_wait'])[:1500])
PY
