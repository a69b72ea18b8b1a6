#!/bin/bash
mkdir -p gpurun_out
./build/xchg > gpurun_out/r02_ubench_xchg.txt 2>&1; cat gpurun_out/r02_ubench_xchg.txt
tools/gpu_variants.sh "base bulk" "p32:1024:65536 p32:2048:65536 p32:4096:32768 p32:8192:16384" "tests/test_gpu_prime.py -k prime32"
tools/gpu_variants.sh "m640 m896" "p64s:1024:65536 p64s:2048:65536 p64s:4096:32768" "tests/test_gpu_prime.py -k prime64"
