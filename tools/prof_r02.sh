#!/bin/bash
# r02: ncu --set full capture of EVERY kernel family of the library (small batches: ncu replays each launch ~40x).  Summaries
# (tools/ncu_summary.py, run on the box) land in gpurun_out/r02_kernels/ncu_<name>.txt; only summaries travel back.
OUT=gpurun_out/r02_kernels; mkdir -p $OUT
run() { # name kernel-regex driver-args...
  local name=$1 rx=$2; shift 2
  if [ -n "$ONLY" ] && ! [[ $name =~ $ONLY ]]; then return; fi   # ONLY=<regex over names>: re-capture a subset
  timeout 900 ncu --set full --clock-control none -k regex:"$rx" -s ${SKIP:-1} -c ${COUNT:-2} -f -o $OUT/prof_$name python tools/prof_driver.py "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py $OUT/prof_$name.ncu-rep > $OUT/ncu_$name.txt 2>&1
  rm -f $OUT/prof_$name.ncu-rep $OUT/ncu_$name.log
}
run ntt64s_2048     'k_ntt_cta'        ntt64 65536 2048          # the bench kernel at the bench grid (roofline.traffic)
run ntt64s_4096     'k_ntt_cta'        ntt64 8192 4096
run ntt64s_65536    'k_ntt_strided'    ntt64 512 65536
run ntt64shoup_2048 'k_ntt_cta'        ntt64shoup 32768 2048
run ntt32_1024      'k_ntt_cta'        ntt32 65536 1024
run ntt32_4096      'k_ntt_cta'        ntt32 16384 4096
run ntt32_8192      'k_ntt_cta'        ntt32 8192 8192
run ntt32_16384     'k_ntt_cta'        ntt32 4096 16384
run ntt32_65536_str 'k_ntt_strided'    ntt32 1024 65536
run ntt32_65536_cta 'k_ntt_cta'        ntt32 1024 65536
run pointwise32     'k_pointwise'      pointwise32 32768 2048
COUNT=3 run pointwise64     'k_pointwise'      pointwise64 32768 2048
run polymul64_2048  'k_polymul_fused'  polymul64 32768 2048
run polymul32_2048  'k_polymul_fused'  polymul32 32768 2048
run polymul128_4096 'k_polymul_fused'  polymul128 2048 4096
run polymulb64_2048 'k_polymul_fused'  polymulb64 32768 2048
SKIP=3 COUNT=3 run polymulb64_65536 'k_large' polymulb64 128 65536
run split64_fwd     'k_native_fwd_fused' split64 16384 2048
SKIP=0 COUNT=3 run split64_inv_crt 'k_native_crt|k_native_reduce' split64inv 8192 2048
SKIP=0 COUNT=2 run plan52 'k_native52' plan52 8192 2048
SKIP=0 COUNT=2 run product_fused 'k_product_fwd_fused|k_product_inv_fused' product 16384 2048
SKIP=0 COUNT=2 run product_generic 'k_product_reduce|k_product_crt' product_generic 8192 2048
SKIP=0 COUNT=1 run product_pointwise 'k_pointwise_strided' product 16384 2048
ls $OUT
