#!/bin/bash
# compute-sanitizer passes over the hot kernels (small batches): memcheck + racecheck + synccheck.  Output: gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out; OUT=gpurun_out/r02_sanitizer.txt; : > $OUT
run() { # tool driver-args
  local tool=$1; shift
  echo "== compute-sanitizer --tool $tool : prof_driver $*" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/prof_driver.py "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|done" | head -12 >> $OUT
}
for t in memcheck racecheck synccheck; do
  run $t ntt32 64 1024
  run $t ntt32 16640 1024      # >= 16384 polynomials: the one-warp-per-polynomial forward kernel (warp-level exchange barrier)
  run $t ntt32 256 256         # groups of <= 32 threads: __syncwarp exchanges
  run $t ntt64 256 256
  run $t polymul64 64 256
  run $t ntt32 8 8192
  run $t ntt32 4 65536
  run $t ntt64 64 2048
  run $t ntt64 4 65536
  run $t polymul64 32 2048
  run $t polymulb64 32 2048
  run $t polymul128 8 4096
  run $t split64inv 16 2048
  run $t plan52 16 2048
  run $t product 32 2048
  run $t product_generic 16 2048
done
if [ -f build/libcntt_cl.so ] && [ -z "$SKIP_CLUSTER" ]; then   # the cluster experiment (DSMEM push exchange), variant library
  export CNTT_B200_LIB=build/libcntt_cl.so
  echo "== variant build/libcntt_cl.so (-DCNTT_CLUSTER32=1)" >> $OUT
  for t in memcheck synccheck racecheck; do run $t ntt32 6 32768; done
  unset CNTT_B200_LIB
fi
cat $OUT
