#!/bin/bash
# r02 round refresh: parity tests, both bench arms, full sweep, launch list of the bench command, configs[0] latency.  Outputs: gpurun_out/r02_*
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 900 python tools/time_polymul.py p32:256:262144 p32:1024:65536 p32:2048:65536 p32:4096:32768 p32:8192:32768 p32:16384:16384 p32:32768:8192 p32:65536:4096 \
  p64s:256:262144 p64s:1024:65536 p64s:2048:65536 p64s:4096:32768 p64s:8192:16384 p64s:16384:8192 p64s:65536:2048 p64:2048:65536 p64:4096:32768 \
  native64:2048:65536 native64:1024:65536 native64:4096:16384 native32:2048:65536 native128:4096:8192 native128:2048:16384 binary64:2048:65536 binary32:2048:65536 binary128:4096:8192 \
  binary64:32768:1024 native64:32768:1024 binary64:65536:1024 pre64:2048:65536 pre128:4096:8192 preb64:2048:65536 preb32:2048:65536 split64:2048:32768 product:2048:65536 > gpurun_out/r02_sweep.txt 2>&1
cat gpurun_out/r02_sweep.txt
timeout 300 python tools/latency_cfg0.py > gpurun_out/r02_latency_cfg0.txt 2>&1; cat gpurun_out/r02_latency_cfg0.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.txt 2>&1; head -30 gpurun_out/r02_launches_bench_summary.txt
