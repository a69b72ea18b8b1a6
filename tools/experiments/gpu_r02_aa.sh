#!/bin/bash
# r02 "warpsync", fourth session: prime32 N = 1024 -- 16 words per thread (64 threads per polynomial, mbinf) against 32 words (one warp, mb0) over the batch size
mkdir -p gpurun_out
OUT=gpurun_out/r02_n10_batch.txt; : > $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
for v in build/libcntt_mbinf.so build/libcntt_mb0.so concrete-ntt_b200/libcntt_b200.so; do
  echo "== $v" | tee -a $OUT
  CNTT_B200_LIB=$v timeout 300 python tools/latency_cfg0.py 2>&1 | tee -a $OUT
  CNTT_B200_LIB=$v timeout 600 python tools/time_polymul.py p32:1024:592 p32:1024:1184 p32:1024:2368 p32:1024:4736 p32:1024:9472 p32:1024:18944 p32:1024:65536 2>&1 | tee -a $OUT
done
