set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py gpu > gpurun_out/r2b_dist.log 2>&1; echo "dist rc=$?"; grep "DIST_GPU_OK\|AssertionError" gpurun_out/r2b_dist.log | head -5
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "bench1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2b_bench1.json','gpurun_out/r2b_bench2.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check'], '%.4g'%d['e2e']['value'])
PY
