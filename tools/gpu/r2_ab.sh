timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 200 > gpurun_out/r2ab_bench8.json 2> gpurun_out/r2ab_bench8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r2ab_bench8.err
python - <<'PY'
import json
for f in ('gpurun_out/r2ab_bench8.json',):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'].get('dist_rel_err'), d['check']['mass_rel_drift'], json.dumps(d['e2e'])[:900]); print(json.dumps(d['halo_wait'])[:1500])
PY
