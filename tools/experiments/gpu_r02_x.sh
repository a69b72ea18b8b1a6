#!/bin/bash
# r02 experiment "warpsync", second session: N = 1024 at 32 words per thread now the default (product plan check), N = 512 the same (n9)
mkdir -p gpurun_out
OUT=gpurun_out/r02_warpsync2.txt; : > $OUT
CNTT_B200_LIB=build/libcntt_n9.so timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_n9.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:512:131072 p32:1024:65536 p32:1024:65537 product:512:131072 product:1024:65536 2>&1 | tee -a $OUT
done
