#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_seq.txt; : > $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_seq.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -1 | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py native64:2048:65536 native64:2048:16384 pre64:2048:65536 preb64:2048:65536 pre128:2048:16384 preb32:2048:65536 2>&1 | tee -a $OUT
done
