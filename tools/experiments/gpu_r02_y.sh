#!/bin/bash
# r02 "warpsync", third session: product plans keep 16 words per thread at N = 1024 (tables of their own)
mkdir -p gpurun_out
OUT=gpurun_out/r02_warpsync3.txt; : > $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a $OUT
timeout 600 python tools/time_polymul.py p32:1024:65536 product:1024:65536 product:512:131072 product:2048:65536 split64:1024:65536 2>&1 | tee -a $OUT
