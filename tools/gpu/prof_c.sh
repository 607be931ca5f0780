# launch list + one full capture of the constant-coefficient stage kernel on the C2 workload
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 4 -c 1 -o gpurun_out/prof_c python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_b2.log 2>&1
ls -la gpurun_out/
