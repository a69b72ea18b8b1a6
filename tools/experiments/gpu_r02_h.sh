#!/bin/bash
mkdir -p gpurun_out
tools/gpu_variants.sh "base nolut" "binary64:2048:65536 binary64:1024:65536 binary64:4096:16384 binary32:2048:65536 binary128:2048:16384 binary128:4096:8192 binary64:256:262144" "tests/test_gpu_native.py tests/test_golden.py tests/test_ref_fixtures.py"
