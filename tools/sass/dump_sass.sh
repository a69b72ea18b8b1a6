#!/bin/bash
# Writes the SASS listing of the headline kernels (profiles/sass/*.sass) and a per-pipe instruction count table.
# usage: tools/sass/dump_sass.sh [library.so]
LIB=${1:-concrete-ntt_b200/libcntt_b200.so}
OUT=profiles/sass; mkdir -p $OUT
cuobjdump -sass $LIB > /tmp/all.sass
dump() { # name mangled-substring
  awk -v pat="$2" '/Function : /{f = index($0, pat) > 0} f' /tmp/all.sass | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//' | grep -v '^\s*/\* 0x' > $OUT/$1.sass
  echo "$1: $(grep -c -E '^\s+/\*[0-9a-f]{4,}\*/' $OUT/$1.sass) instructions"
}
dump ntt32_n1024_fwd   'k_ntt_ctaINS_5A32L4ELi10ELi4ELi2ELb1ELb1ELi1'
dump ntt32_n1024_fwd_pipe 'k_ntt_cta_pipeINS_5A32L4ELi10ELi4ELi2ELb1ELi1'
dump ntt32_n1024_inv   'k_ntt_ctaINS_5A32L4ELi10ELi4ELi2ELb0ELb1ELi1'
dump ntt64s_n2048_fwd  'k_ntt_ctaINS_4A64SELi11ELi4ELi1ELb1ELb1ELi1'
dump ntt64s_n2048_inv  'k_ntt_ctaINS_4A64SELi11ELi4ELi1ELb0ELb1ELi1'
dump ntt64l4_n2048_fwd 'k_ntt_ctaINS_5A64L4ELi11ELi4ELi1ELb1ELb1ELi1'
dump strided32_k4_fwd  'k_ntt_stridedINS_5A32L4ELi4ELb1'
dump pointwise32_mul_assign_normalize 'k_pointwiseINS_5A32L4ELi0'
dump polymul_native64_n2048 'k_polymul_fusedILi1ELi11ELi4'
dump polymul_native128_n4096 'k_polymul_fusedILi2ELi12ELi3'
dump large_binary64_lead_fwd_c16 'k_large_lead_fwdILi4ELi4E'
dump large_mid_c16 'k_large_midILi4E'
dump large_binary64_lead_inv_c16 'k_large_lead_invILi4ELi4E'
python tools/sass/pipecount.py $LIB 'k_ntt_ctaINS_5A32L4ELi10ELi4ELi2' > $OUT/pipecount.txt
python tools/sass/pipecount.py $LIB 'k_ntt_ctaINS_4A64SELi11ELi4ELi1' >> $OUT/pipecount.txt
python tools/sass/pipecount.py $LIB 'k_polymul_fusedILi1ELi11ELi4' >> $OUT/pipecount.txt
