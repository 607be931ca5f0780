for m in 1 2; do
RMH_DEBUG_HALO=$m RMH_NO_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2958$m bench.py --gpus 2 --steps 60 --no-dist-check > gpurun_out/r2y_dbg$m.json 2> gpurun_out/r2y_dbg$m.err; echo "rc=$?"
done
RMH_NO_FUSED_SEND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29585 bench.py --gpus 2 --steps 60 --no-dist-check > gpurun_out/r2y_dbg0.json 2> gpurun_out/r2y_dbg0.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2y_dbg1.json','gpurun_out/r2y_dbg2.json','gpurun_out/r2y_dbg0.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('gpu_launches'), json.dumps(d['halo_wait']['per_rank']))
PY
