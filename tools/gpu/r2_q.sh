set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_worker.py gpu > gpurun_out/r2q_worker.log 2>&1; echo "worker rc=$?"; tail -12 gpurun_out/r2q_worker.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 200 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2q_bench2.err
RMH_NO_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 200 --no-dist-check > gpurun_out/r2q_bench2_nosend.json 2> gpurun_out/r2q_bench2_nosend.err; echo "bench2 nosend rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2q_bench2.json','gpurun_out/r2q_bench2_nosend.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'].get('dist_rel_err'), d['check']['mass_rel_drift'], d.get('gpu_launches'))
PY
