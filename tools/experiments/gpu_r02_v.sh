#!/bin/bash
# r02 experiment "np2sub": sub-block launches of multi-launch 32-bit transforms carry the same sub-block of two polynomials per group
mkdir -p gpurun_out
OUT=gpurun_out/r02_np2sub.txt; : > $OUT
CNTT_B200_LIB=build/libcntt_np2.so timeout 900 python -m pytest tests/test_gpu_prime.py tests/test_gpu_product.py tests/test_gpu_native.py -m gpu -q -x -k "large or two_level or product or 32768 or 65536 or persistent or full_size" 2>&1 | tail -2 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_np2.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:32768:8192 p32:65536:4096 p32:65536:333 p32:131072:2048 binary64:32768:1024 native64:32768:1024 binary64:65536:1024 2>&1 | grep -v Traceback | grep "n=" | tee -a $OUT
done
