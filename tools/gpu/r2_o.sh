set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 200 > gpurun_out/r2o_bench8.json 2> gpurun_out/r2o_bench8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r2o_bench8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --order 4 --nloc 74 --steps 200 --no-dist-check > gpurun_out/r2o_bench8_c4.json 2> gpurun_out/r2o_bench8_c4.err; echo "bench8 c4 rc=$?"
RMH_SLOW_TESTS=1 timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q > gpurun_out/r2o_scale.log 2>&1; tail -3 gpurun_out/r2o_scale.log
python - <<'PY'
import json
for f in ('gpurun_out/r2o_bench8.json','gpurun_out/r2o_bench8_c4.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'].get('dist_rel_err'), d['check']['mass_rel_drift'], '%.4g'%d['e2e']['value'])
PY
