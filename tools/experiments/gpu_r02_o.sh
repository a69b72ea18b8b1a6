#!/bin/bash
# r02 experiment "l2chunk": multi-launch transforms walked in L2-sized slices of the batch (CNTT_L2_CHUNK_MB, ntt_kernels.cuh)
mkdir -p gpurun_out
OUT=gpurun_out/r02_l2chunk.txt; : > $OUT
CNTT_L2_CHUNK_MB=16 timeout 900 python -m pytest tests/test_gpu_prime.py -m gpu -q -k "large or two_level or full_size" 2>&1 | tail -2 | tee -a $OUT
for c in 0 8 16 24 32 48 64; do
  echo "== CNTT_L2_CHUNK_MB=$c" | tee -a $OUT
  CNTT_L2_CHUNK_MB=$c timeout 600 python tools/time_polymul.py p32:32768:8192 p32:65536:4096 p32:131072:2048 p64s:8192:16384 p64s:16384:8192 p64s:65536:2048 p64:8192:16384 2>&1 | tee -a $OUT
done
