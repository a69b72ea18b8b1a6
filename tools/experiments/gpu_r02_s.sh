#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_acc_occ2.txt; : > $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_a512.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py native32:1024:65536 native32:4096:16384 binary32:2048:65536 binary32:1024:65536 binary128:4096:8192 binary128:2048:16384 native128:1024:32768 pre128:4096:8192 pre128:2048:16384 pre32:2048:65536 pre64:1024:65536 pre64:4096:16384 preb64:2048:65536 preb32:2048:65536 2>&1 | tee -a $OUT
done
