#!/bin/bash
# ncu --set full captures of the main kernels (small batches keep the replays short). Outputs under gpurun_out/.
mkdir -p gpurun_out
for w in "ntt32 65536 1024" "ntt64 65536 2048" "polymul64 32768 2048"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_cta|k_polymul_fused' -s 2 -c 2 -f -o gpurun_out/prof_$1 python tools/prof_driver.py $1 $2 $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
done
ls -la gpurun_out
