# two-rank weak-scaling bench through the driver's launch line (NCCL halo exchange per stage)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --problem 1 > gpurun_out/bench_2gpu_p1.json 2> gpurun_out/bench_2gpu_p1.err
cat gpurun_out/bench_2gpu_p1.json; tail -3 gpurun_out/bench_2gpu_p1.err
