set -x
timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q -k "decomposed" > gpurun_out/r2c_cli.log 2>&1; echo "cli rc=$?"; tail -30 gpurun_out/r2c_cli.log
timeout 300 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 remhos_b200/host/remhos -m tests/data/periodic-cube.mesh -p 0 -rs 3 -o 3 -dt 0.002 -tf 0.1 -ho 3 -lo 5 -fct 2 -pa -no-vis > gpurun_out/r2c_torchrun_cli.log 2>&1; echo "torchrun cli rc=$?"; tail -12 gpurun_out/r2c_torchrun_cli.log
for P in 0 1; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_stage3 -s 6 -c 6 --csv --log-file gpurun_out/traffic_o3_p$P.csv python bench.py --problem $P --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_stage3 -s 6 -c 6 --csv --log-file gpurun_out/traffic_o4_p0.csv python bench.py --order 4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --order 4 --steps 50 > gpurun_out/r2c_bench2_o4.json 2> gpurun_out/r2c_bench2_o4.err; echo "bench2 o4 rc=$?"; tail -c 1500 gpurun_out/r2c_bench2_o4.json
