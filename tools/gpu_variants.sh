#!/bin/bash
# A/B session: for each experimental library build/libcntt_<v>.so run parity tests and a timing sweep.
# usage: tools/gpu_variants.sh "v1 v2 ..." "case case ..." ["test files"]     (outputs under gpurun_out/)
mkdir -p gpurun_out
TESTS=${3:-tests/test_gpu_prime.py}
for v in $1; do
  export CNTT_B200_LIB=$PWD/build/libcntt_$v.so
  [ "$v" = base ] && export CNTT_B200_LIB=$PWD/concrete-ntt_b200/libcntt_b200.so
  timeout 600 python -m pytest $TESTS -m gpu -x -q > gpurun_out/var_${v}_pytest.log 2>&1; echo "$v pytest rc=$? $(tail -1 gpurun_out/var_${v}_pytest.log)"
  timeout 300 python tools/time_polymul.py $2 > gpurun_out/var_${v}_sweep.txt 2>&1
  sed "s/^/$v: /" gpurun_out/var_${v}_sweep.txt
done
