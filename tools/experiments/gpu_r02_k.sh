#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base" "native64:2048:65536 pre64:2048:65536 pre64:1024:65536 pre32:2048:65536 pre128:4096:8192 native128:4096:8192" "tests/test_gpu_native.py"
