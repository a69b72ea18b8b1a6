#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base" "p64s:1024:65536 p64s:2048:65536 p64s:4096:32768" "tests/test_gpu_prime.py -k prime64"
