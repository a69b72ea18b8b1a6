#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base rl rlm" "native64:2048:65536 native64:1024:65536 native64:4096:16384 binary64:2048:65536 native32:2048:65536" "tests/test_gpu_native.py -k polymul_matches"
