bash tools/prof_all.sh > gpurun_out/prof_all.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v4.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
tail -25 gpurun_out/round.log
