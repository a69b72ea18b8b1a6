#!/bin/bash
# One GPU-box session: parity tests, bench, quick timing sweep, ncu launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python tools/time_polymul.py p32:1024:65536 p32:2048:65536 p32:4096:32768 p64s:2048:65536 p64:2048:65536 p64s:1024:65536 p64s:4096:32768 p64s:16384:8192 p64s:65536:2048 p32:65536:4096 native64:2048:65536 native64:1024:65536 native64:4096:16384 native32:2048:65536 native128:4096:8192 binary64:2048:65536 binary64:32768:1024 native64:32768:1024 > gpurun_out/sweep.txt 2>&1
cat gpurun_out/sweep.txt
