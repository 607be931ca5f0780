set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 100 --no-dist-check > gpurun_out/r2z_bench2.json 2> gpurun_out/r2z_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2z_bench2.err
RMH_NO_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 100 --no-dist-check > gpurun_out/r2z_bench2_nosend.json 2> gpurun_out/r2z_bench2_nosend.err; echo "bench2 nosend rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2z_bench2.json','gpurun_out/r2z_bench2_nosend.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d.get('gpu_launches'), json.dumps(d['halo_wait']['per_rank']), d['e2e']['value'], d['e2e']['serial_value'], d['e2e']['independent_fields_value'], d['e2e']['pipelined_equals_serial'])
PY
