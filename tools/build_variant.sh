#!/bin/bash
# Builds an experimental variant of the library: tools/build_variant.sh NAME "<extra nvcc flags>"  ->  build/libcntt_NAME.so
# (load it with CNTT_B200_LIB=build/libcntt_NAME.so; experiments only, never shipped)
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../concrete-ntt_b200/csrc"
OBJ=../../build/obj_$NAME; mkdir -p $OBJ
for f in capi native_kernels native_split product_fused inst_a32l4 inst_a32l2 inst_a32g inst_a64l4 inst_a64l2 inst_a64s inst_a64g; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $EXTRA -c $f.cu -o $OBJ/$f.o 2> $OBJ/$f.ptxas.log || { cat $OBJ/$f.ptxas.log; exit 1; } ) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --no-undefined -o ../../build/libcntt_$NAME.so $OBJ/*.o
echo built build/libcntt_$NAME.so
