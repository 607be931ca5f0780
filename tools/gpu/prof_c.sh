# launch list + one full capture of the constant-coefficient stage kernel on the C2 workload
ncu --set full --clock-control none --import-source on -k regex:k_stage3c -s 4 -c 1 -o gpurun_out/prof_c python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_b2.log 2>&1
ncu --set full --clock-control none -k regex:k_ent_min_max -s 4 -c 1 -o gpurun_out/prof_ent python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_b3.log 2>&1
python bench.py --problem 1 --no-cpu-baseline > gpurun_out/bench_p1.json 2>/dev/null
python bench.py --order 4 --rs 5 --steps 10 --no-cpu-baseline > gpurun_out/bench_o4.json 2>/dev/null
ls -la gpurun_out/
