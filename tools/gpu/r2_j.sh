set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2_remap.csv python tools/bench_remap.py 5 2 > gpurun_out/r2j_remap_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_geom3 -s 4 -c 1 -o gpurun_out/prof_r2_geom3 python tools/bench_remap.py 5 2 > gpurun_out/r2j_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 4 -c 1 -o gpurun_out/prof_r2_kstage python tools/bench_remap.py 5 2 > gpurun_out/r2j_b.log 2>&1
python tools/bench_remap.py 5 5
# 2-GPU recheck of the reshaped put kernel
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_worker.py gpu 2>&1 | grep "DIST_GPU_OK\|Error"
