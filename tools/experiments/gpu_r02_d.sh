#!/bin/bash
mkdir -p gpurun_out
CASES="p64s:1024:65536 p64s:2048:65536 p64s:4096:32768 p64s:8192:16384 p64s:256:262144 p64s:512:131072 p64s:65536:2048"
tools/gpu_variants.sh "base noff noshift" "$CASES" "tests/test_gpu_prime.py -k prime64"
