#!/bin/bash
# r02 experiment "w1112": forward transforms of N = 2048 / 4096 x 32-bit (large batches) at 32 words per thread (-DCNTT_R32_WHOLE_MASK=0x1800)
mkdir -p gpurun_out
OUT=gpurun_out/r02_w1112.txt; : > $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
CNTT_B200_LIB=build/libcntt_w1112.so timeout 1200 python -m pytest tests/test_gpu_prime.py -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_w1112.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:2048:65536 p32:2048:20000 p32:4096:32768 p32:4096:16384 2>&1 | tee -a $OUT
done
