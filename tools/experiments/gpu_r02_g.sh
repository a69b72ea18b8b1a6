#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base" "p64s:8192:16384 p64s:16384:8192 p64s:32768:4096 p64s:65536:2048" "tests/test_gpu_prime.py"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
CNTT_B200_LIB=$PWD/build/libcntt_bulk.so timeout 600 python -m pytest tests/test_gpu_prime.py -m gpu -x -q -k "prime32" 2>&1 | tail -2
