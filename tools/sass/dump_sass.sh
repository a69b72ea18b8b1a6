#!/bin/bash
# Writes one SASS listing per kernel family of the library (profiles/sass/*.sass: the instantiation the bench / BASELINE configs
# launch, or the most used one) and a per-pipe static instruction count table.   usage: tools/sass/dump_sass.sh [library.so]
LIB=${1:-concrete-ntt_b200/libcntt_b200.so}
OUT=profiles/sass; mkdir -p $OUT; rm -f $OUT/*.sass
cuobjdump -sass $LIB > /tmp/all.sass
dump() { # name mangled-substring
  awk -v pat="$2" '/Function : /{f = index($0, pat) > 0} f' /tmp/all.sass | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//' | grep -v '^\s*/\* 0x' > $OUT/$1.sass
  n=$(grep -c -E '^\s+/\*[0-9a-f]{4,}\*/' $OUT/$1.sass)
  [ "$n" -gt 0 ] || { echo "$1: NO SUCH KERNEL ($2)"; rm -f $OUT/$1.sass; return; }
  echo "$1: $n instructions"
}
# single-prime transforms (k_ntt_cta): bench headline (Solinas N=2048), BASELINE configs[0] kernel (prime32 N=1024), the 32-words-per-
# thread kernels of N = 8192, a Shoup-64 class, a sub-block flavour; strided leading levels; table builder
dump ntt64s_n2048_fwd  'k_ntt_ctaINS_4A64SELi11ELi4ELi1ELb1ELb1ELi1'
dump ntt64s_n2048_inv  'k_ntt_ctaINS_4A64SELi11ELi4ELi1ELb0ELb1ELi1'
dump ntt32_n1024_fwd   'k_ntt_ctaINS_5A32L4ELi10ELi5ELi4ELb1ELb1ELi1'
dump ntt32_n1024_inv   'k_ntt_ctaINS_5A32L4ELi10ELi5ELi4ELb0ELb1ELi1'
dump ntt32_n8192_fwd_r32 'k_ntt_ctaINS_5A32L4ELi13ELi5ELi1ELb1ELb1ELi1'
dump ntt32_n8192_inv_r32 'k_ntt_ctaINS_5A32L4ELi13ELi5ELi1ELb0ELb1ELi1'
dump ntt64l4_n2048_fwd 'k_ntt_ctaINS_5A64L4ELi11ELi4ELi1ELb1ELb1ELi1'
dump ntt32_sub4096_fwd 'k_ntt_ctaINS_5A32L4ELi12ELi4ELi1ELb1ELb0ELi1'
dump strided32_k4_fwd  'k_ntt_stridedINS_5A32L4ELi4ELb1'
dump strided64s_k4_inv 'k_ntt_stridedINS_4A64SELi4ELb0'
dump build_last32_n1024 'k_build_lastINS_6EngineINS_5A32L4ELi10ELi5'
# pointwise streams
dump pointwise32_mul_assign_normalize 'k_pointwiseINS_5A32L4ELi0'
dump pointwise64s_mul_accumulate 'k_pointwiseINS_4A64SELi2'
dump pointwise32_strided_mul_accumulate 'k_pointwise_stridedINS_5A32L4ELi2'
# native plans: fused polymul (configs[2], configs[3]), split fwd, reduce / crt, large-N pipeline (configs[4]), Plan52
dump polymul_native64_n2048 'k_polymul_fusedILi1ELi11ELi4ELb0'
dump polymul_native64_pre_n2048 'k_polymul_fusedILi1ELi11ELi4ELb1'   # rhs already in the NTT domain (register CRT accumulation, four CTAs per SM)
dump polymul_native128_n4096 'k_polymul_fusedILi2ELi12ELi3'
dump native64_fwd_fused_n2048 'k_native_fwd_fusedILi1ELi11ELb0'
dump native64_reduce 'k_native_reduceILi8ELi5ELb0'
dump native64_crt 'k_native_crtILi1E'
dump native128_crt 'k_native_crtILi2E'
dump large_binary64_lead_fwd_c16 'k_large_lead_fwdILi4ELi4E'
dump large_mid_c16 'k_large_midILi4E'
dump large_binary64_lead_inv_c16 'k_large_lead_invILi4ELi4E'
dump native52_reduce 'k_native52_reduceILb0'
dump native52_crt 'k_native52_crt'
# product::Plan
dump product_fwd_fused_n2048 'k_product_fwd_fusedINS_5A32L2ELi11'
dump product_inv_fused_n2048 'k_product_inv_fusedINS_5A32L2ELi11'
dump product_reduce 'k_product_reduce'
dump product_crt 'k_product_crt'
: > $OUT/pipecount.txt
for pat in 'k_ntt_ctaINS_4A64SELi11ELi4ELi1' 'k_ntt_ctaINS_5A32L4ELi10ELi5ELi4' 'k_ntt_ctaINS_5A32L4ELi13ELi5ELi1ELb' 'k_polymul_fusedILi1ELi11ELi4' 'k_polymul_fusedILi2ELi12ELi3' 'k_native_fwd_fusedILi1ELi11ELb0' 'k_product_fwd_fusedINS_5A32L2ELi11' 'k_product_inv_fusedINS_5A32L2ELi11'; do
  python tools/sass/pipecount.py $LIB "$pat" >> $OUT/pipecount.txt
done
