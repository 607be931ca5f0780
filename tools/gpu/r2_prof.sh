# launch list + one full capture of the folded constant-coefficient stage kernel on the C2 workload (1 GPU)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_fold_rs5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_r2_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 6 -c 1 -o gpurun_out/prof_r2_fold python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_r2_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 6 -c 1 -o gpurun_out/prof_r2_fold_o4 python bench.py --order 4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_r2_b.log 2>&1
python bench.py --order 4 --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_o4.json 2>/dev/null
ls -la gpurun_out/
