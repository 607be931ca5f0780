set -x
RMH_FORCE_GH=1 timeout 500 python bench.py --force-dist --nloc 96 --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r2t_fd.json 2> gpurun_out/r2t_fd.err; echo rc=$?; tail -3 gpurun_out/r2t_fd.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2t_fd.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'])"
RMH_FORCE_GH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 6 -c 1 -o gpurun_out/prof_r2t_gh python bench.py --force-dist --nloc 96 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_r2t.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 6 -c 1 -o gpurun_out/prof_r2t_nogh python bench.py --force-dist --nloc 96 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_r2t2.log 2>&1
ls -la gpurun_out/prof_r2t_*
