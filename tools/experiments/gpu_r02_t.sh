#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_acc_occ3.txt; : > $OUT
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
timeout 600 python tools/time_polymul.py native64:2048:65536 native64:1024:65536 native32:2048:65536 native32:1024:65536 binary32:2048:65536 binary64:2048:65536 native128:4096:8192 native128:2048:16384 native128:1024:32768 native128:256:65536 binary128:2048:16384 pre128:2048:16384 pre32:2048:65536 pre64:2048:65536 pre64:1024:65536 preb64:2048:65536 preb32:2048:65536 2>&1 | tee -a $OUT
