#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base k64s4" "p64s:8192:16384 p64s:16384:8192 p64s:32768:4096 p64s:65536:2048 p64s:131072:1024" "tests/test_gpu_prime.py -k large_n"
