RMH_FORCE_GH=1 timeout 500 python bench.py --force-dist --nloc 96 --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r2aa_fd.json 2> gpurun_out/r2aa_fd.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_worker.py gpu > gpurun_out/r2aa_worker.log 2>&1; echo "worker rc=$?"; tail -1 gpurun_out/r2aa_worker.log
RMH_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29574 tests/dist_worker.py gpu > gpurun_out/r2aa_worker_send.log 2>&1; echo "worker send rc=$?"; tail -1 gpurun_out/r2aa_worker_send.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 200 > gpurun_out/r2aa_bench2.json 2> gpurun_out/r2aa_bench2.err; echo "bench2 rc=$?"
RMH_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 100 --no-dist-check > gpurun_out/r2aa_bench2_send.json 2> gpurun_out/r2aa_bench2_send.err; echo "bench2 send rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2aa_fd.json','gpurun_out/r2aa_bench2.json','gpurun_out/r2aa_bench2_send.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d.get('gpu_launches'), d['check'].get('dist_rel_err'), json.dumps((d.get('halo_wait') or {}).get('per_rank')), d['e2e']['value'], d['e2e']['serial_value'], d['e2e']['independent_fields_value'])
PY
