timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_worker.py gpu > gpurun_out/r2z_worker.log 2>&1; echo "worker rc=$?"; tail -2 gpurun_out/r2z_worker.log
bash tools/gpu/r2_u.sh
