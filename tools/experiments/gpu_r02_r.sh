#!/bin/bash
# r02 experiment "acc occupancy": fused polymul with register CRT accumulation + per-prime operand reload for every kind, compiled for
# 384 / 512 / 640 resident threads (3 / 4 / 5 CTAs of 128 threads at N = 2048); and the graph-capture test
mkdir -p gpurun_out
OUT=gpurun_out/r02_acc_occ.txt; : > $OUT
timeout 600 python -m pytest tests/test_gpu_prime.py -m gpu -q -k "graph_capturable" 2>&1 | tail -3 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_a384.so build/libcntt_a512.so build/libcntt_a640.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py native64:2048:65536 native64:1024:65536 native64:4096:16384 native32:2048:65536 binary64:2048:65536 native128:4096:8192 native128:2048:16384 pre64:2048:65536 2>&1 | tee -a $OUT
done
CNTT_B200_LIB=build/libcntt_a512.so timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
