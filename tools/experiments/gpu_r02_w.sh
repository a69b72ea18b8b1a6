#!/bin/bash
# r02 experiment "warpsync": exchange barrier of groups of <= 32 threads as __syncwarp (base) vs __syncthreads (ws0), and N = 1024 x u32 with
# 32 words per thread = one warp per polynomial, one exchange (n10)
mkdir -p gpurun_out
OUT=gpurun_out/r02_warpsync.txt; : > $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_n10.so; do
  echo "== tests $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
done
for v in build/libcntt_ws0.so concrete-ntt_b200/libcntt_b200.so build/libcntt_n10.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:64:1048576 p32:128:524288 p32:256:262144 p32:512:131072 p32:1024:65536 p64s:64:524288 p64s:128:524288 p64s:256:262144 p64s:512:131072 p64:256:262144 native64:256:262144 native64:512:131072 binary64:256:262144 native32:256:262144 native128:256:65536 product:256:262144 product:512:131072 2>&1 | tee -a $OUT
done
