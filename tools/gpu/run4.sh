python -m pytest tests/test_gpu_stage.py -x -q > gpurun_out/t4.log 2>&1; tail -3 gpurun_out/t4.log
P='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["roofline"]["kernel"][:40], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], d["check"], d["e2e"]["value"], d["config"]["dofs_per_gpu"])
    elif "k_stage3" in l or "k_op_linear" in l: print(l.strip())
'
echo "== default"; RMH_VERBOSE=1 python bench.py --steps 10 --no-cpu-baseline 2>&1 | python -c "$P"
echo "== problem 1"; RMH_VERBOSE=1 python bench.py --steps 10 --no-cpu-baseline --problem 1 2>&1 | python -c "$P"
echo "== order 4 rs 5"; RMH_VERBOSE=1 python bench.py --steps 6 --no-cpu-baseline --order 4 --rs 5 2>&1 | python -c "$P"
echo "== order 2 rs 5"; RMH_VERBOSE=1 python bench.py --steps 10 --no-cpu-baseline --order 2 --rs 5 2>&1 | python -c "$P"
