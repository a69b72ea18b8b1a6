#!/bin/bash
# r02 session B: butterfly micro-benchmark (fixed host check), LOGR=5 variants, ncu of the one-CTA N=32768 kernel
mkdir -p gpurun_out
./build/gold_bf > gpurun_out/r02_gold_bf.txt 2>&1; echo "gold_bf rc=$?"; cat gpurun_out/r02_gold_bf.txt
CASES="p32:8192:32768 p32:16384:16384 p32:32768:8192 p32:65536:4096"
tools/gpu_variants.sh "base r32 r32n1 r32n1np" "$CASES" "tests/test_gpu_prime.py -k large_n"
for n in 8192 32768; do
  CNTT_B200_LIB=$PWD/build/libcntt_r32n1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_cta' -s 2 -c 2 -f -o gpurun_out/prof_r32n1_$n python tools/prof_driver.py ntt32 $((67108864 / n)) $n > gpurun_out/ncu_r32n1_$n.log 2>&1
  echo "ncu $n rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_r32n1_$n.ncu-rep > gpurun_out/ncu_r32n1_$n.txt 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
