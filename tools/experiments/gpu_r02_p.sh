#!/bin/bash
# r02 experiment "cluster": N = 32768 (32-bit words) as one launch of two-CTA clusters, DSMEM exchange after pass 0 (-DCNTT_CLUSTER32=1)
mkdir -p gpurun_out
OUT=gpurun_out/r02_cluster.txt; : > $OUT
CNTT_B200_LIB=build/libcntt_cl.so timeout 600 python -m pytest tests/test_gpu_prime.py tests/test_gpu_product.py -m gpu -q -k "large_n_two_level or product" 2>&1 | tail -3 | tee -a $OUT
for v in concrete-ntt_b200/libcntt_b200.so build/libcntt_cl.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:32768:8192 p32:32768:1024 p32:32768:300 2>&1 | tee -a $OUT
done
