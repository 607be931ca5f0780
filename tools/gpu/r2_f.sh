set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tests/dist_worker.py gpu > gpurun_out/r2f_dist4.log 2>&1; echo "dist4 rc=$?"; grep "DIST_GPU_OK\|Error" gpurun_out/r2f_dist4.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --rs 4 --steps 100 > gpurun_out/r2f_bench4_rs4.json 2> gpurun_out/r2f_bench4_rs4.err; echo "bench4 rc=$?"; tail -c 900 gpurun_out/r2f_bench4_rs4.json; tail -3 gpurun_out/r2f_bench4_rs4.err
timeout 300 remhos_b200/host/remhos -gpus 4 -m tests/data/periodic-cube.mesh -p 0 -rs 3 -o 3 -dt 0.002 -tf 0.1 -ho 3 -lo 5 -fct 2 -pa -no-vis | tail -8
