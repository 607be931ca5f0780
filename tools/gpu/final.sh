# end-of-session validation: smoke, full GPU suite, default bench (N=1), launch list
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t_final.log 2>&1; tail -4 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/launches_v13.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
