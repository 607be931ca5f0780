# round 2, first GPU pass (2 GPUs): dist parity test, whole GPU suite, bench at N=1 and N=2
set -x
nvidia-smi -L
export RMH_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py gpu > gpurun_out/r2a_dist.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/r2a_dist.log
unset RMH_VERBOSE
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 50 > gpurun_out/r2a_bench1.json 2> gpurun_out/r2a_bench1.err; echo "bench1 rc=$?"; cat gpurun_out/r2a_bench1.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 > gpurun_out/r2a_bench2.json 2> gpurun_out/r2a_bench2.err; echo "bench2 rc=$?"; cat gpurun_out/r2a_bench2.json | cut -c1-1500; tail -3 gpurun_out/r2a_bench2.err
