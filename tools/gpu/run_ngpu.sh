# N-rank weak-scaling bench through the driver's launch line (NCCL halo exchange per stage)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_${N}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel'][:30], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'], d['gpu_launches'], d['e2e']['value'])"
wc -l gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
