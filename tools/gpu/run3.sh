python -m pytest tests/test_gpu_stage.py -x -q > gpurun_out/t3.log 2>&1; tail -5 gpurun_out/t3.log
for cfg in 0 822 1022; do echo "== C_CFG=$cfg"; RMH_VERBOSE=1 RMH_C_CFG=$cfg python bench.py --steps 10 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['check'], d['e2e']['value'])
    elif 'k_stage3' in l or 'k_op_linear' in l: print(l.strip())
"; done
bash tools/gpu/prof_c.sh > /dev/null 2>&1
