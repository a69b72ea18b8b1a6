#!/bin/bash
# r02 session A: Goldilocks butterfly micro-benchmark, LOGR=5 large-N A/B for 32-bit words, full GPU test suite, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
./build/gold_bf > gpurun_out/r02_gold_bf.txt 2>&1; echo "gold_bf rc=$?"; cat gpurun_out/r02_gold_bf.txt
CASES="p32:4096:32768 p32:8192:32768 p32:16384:16384 p32:32768:8192 p32:65536:4096 p32:131072:2048"
tools/gpu_variants.sh "base r32off r32nopipe r32blk14" "$CASES" "tests/test_gpu_prime.py -k large_n"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
