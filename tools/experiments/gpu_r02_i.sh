#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base" "p32:32768:4096 p32:65536:2048 p32:131072:1024" "tests/test_gpu_prime.py -k large_n"
