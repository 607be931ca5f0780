set -x
# A: two independent single-GPU benches at the same time, no process group at all
(CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --nloc 96 --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r2w_a0.json 2> gpurun_out/r2w_a0.err) &
(CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --nloc 96 --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r2w_a1.json 2> gpurun_out/r2w_a1.err) &
wait
# B: the same under torchrun with the NCCL process group up, ghost-aware kernel, no peers
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --replicas --force-dist --nloc 96 --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r2w_b.json 2> gpurun_out/r2w_b.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r2w_a0.json','gpurun_out/r2w_a1.json','gpurun_out/r2w_b.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['serial_value'])
PY
tail -3 gpurun_out/r2w_b.err
