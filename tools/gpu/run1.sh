python -m pytest tests/test_gpu_stage.py -x -q -k linear_operator > gpurun_out/t1.log 2>&1; tail -5 gpurun_out/t1.log
for cfg in "8 2" "10 2" "5 4" "4 4" "7 2" "6 2"; do set -- $cfg; echo "== NW=$1 MINB=$2"; RMH_VERBOSE=1 RMH_W_NW=$1 RMH_W_MINB=$2 python bench.py --steps 10 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['check'])
    elif 'k_stage3w' in l: print(l.strip())
"; done
echo "== stored"; RMH_NO_LINEAR_OP=1 python bench.py --steps 10 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['check'])"
