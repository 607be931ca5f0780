set -x
timeout 900 python -m pytest tests/test_cli.py -m gpu -q -k "decomposed" > gpurun_out/r2m_cli.log 2>&1; echo "cli rc=$?"; tail -40 gpurun_out/r2m_cli.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 100 > gpurun_out/r2m_bench2.json 2> gpurun_out/r2m_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2m_bench2.json',):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['check']['dist_rel_err'], '%.4g'%d['e2e']['value'])
PY
