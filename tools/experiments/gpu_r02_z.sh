#!/bin/bash
# r02 experiment "n11": N = 2048 x 32-bit with 64 words per thread (one warp per polynomial, 5 + 6 levels, one exchange)
mkdir -p gpurun_out
OUT=gpurun_out/r02_n11.txt; : > $OUT
CNTT_B200_LIB=build/libcntt_n11.so timeout 1200 python -m pytest tests/test_gpu_prime.py tests/test_gpu_product.py -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_n11.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:2048:65536 p32:65536:4096 product:2048:65536 2>&1 | tee -a $OUT
done
