python -m pytest tests -m gpu -x -q > gpurun_out/t5.log 2>&1; tail -3 gpurun_out/t5.log
python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json; tail -2 gpurun_out/bench5.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/launches_v12.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
