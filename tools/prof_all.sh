#!/bin/bash
# ncu --set full capture of every kernel family of the library (small batches: ncu replays each launch ~40x).
# Summaries (tools/ncu_summary.py, run on the box) land in gpurun_out/ncu_sum_<name>.txt; the .ncu-rep files of the
# three headline kernels come from tools/gpu_ncu.sh (with --import-source on).
mkdir -p gpurun_out
run() { # name kernel-regex driver-args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none -k regex:"$rx" -s ${SKIP:-1} -c ${COUNT:-2} -f -o gpurun_out/prof_$name python tools/prof_driver.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_sum_$name.txt 2>&1
  rm -f gpurun_out/prof_$name.ncu-rep gpurun_out/ncu_$name.log   # the 64 MiB return limit: only summaries travel back
}
run ntt32_1024      'k_ntt_cta'        ntt32 65536 1024
run ntt32_4096      'k_ntt_cta'        ntt32 16384 4096
run ntt32_65536_str 'k_ntt_strided'    ntt32 1024 65536
run ntt32_65536_cta 'k_ntt_cta'        ntt32 1024 65536
run ntt64s_2048     'k_ntt_cta'        ntt64 65536 2048
run ntt64s_65536    'k_ntt_strided'    ntt64 512 65536
run ntt64shoup_2048 'k_ntt_cta'        ntt64shoup 32768 2048
run pointwise32     'k_pointwise'      pointwise32 32768 2048
run pointwise64     'k_pointwise'      pointwise64 32768 2048
run polymul64_2048  'k_polymul_fused'  polymul64 32768 2048
run polymul32_2048  'k_polymul_fused'  polymul32 32768 2048
run polymul128_4096 'k_polymul_fused'  polymul128 2048 4096
run polymulb64_2048 'k_polymul_fused'  polymulb64 32768 2048
SKIP=3 COUNT=3 run polymulb64_32768 'k_large' polymulb64 256 32768
SKIP=3 COUNT=3 run polymulb64_65536 'k_large' polymulb64 128 65536
ls -la gpurun_out/
