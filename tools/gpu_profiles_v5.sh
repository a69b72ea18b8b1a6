#!/bin/bash
# Final evidence pass of the round: ncu summaries of every kernel family, stall / opcode breakdown of the headline kernels,
# launch list of the bench command, full parity suite, bench line, sweep.  Outputs under gpurun_out/.
mkdir -p gpurun_out
bash tools/prof_all.sh > gpurun_out/prof_all.log 2>&1
SKIP=3 COUNT=3 bash -c 'source /dev/null; name=product_2048; timeout 600 ncu --set full --clock-control none -k regex:"k_product" -s 2 -c 2 -f -o gpurun_out/prof_$name python tools/time_polymul.py product:2048:32768 > gpurun_out/ncu_$name.log 2>&1; python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_sum_$name.txt 2>&1; rm -f gpurun_out/prof_$name.ncu-rep'
for k in "pm64 k_polymul_fused polymul64 32768 2048" "n64s k_ntt_cta ntt64 65536 2048" "n32inv k_ntt_cta< ntt32 65536 1024" "n32fwd k_ntt_cta_pipe ntt32 65536 1024"; do
  set -- $k; bash tools/gpu_stalls.sh "$@" > /dev/null 2>&1; rm -f gpurun_out/st_$1.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v5.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; cat gpurun_out/bench_v5.json | cut -c1-300
timeout 600 python tools/time_polymul.py p32:256:262144 p32:1024:65536 p32:2048:65536 p32:4096:32768 p32:16384:16384 p32:65536:4096 p64s:1024:65536 p64s:2048:65536 p64:2048:65536 p64s:4096:32768 p64s:16384:8192 p64s:65536:2048 native64:1024:65536 native64:2048:65536 native64:4096:16384 native32:2048:65536 native128:2048:16384 native128:4096:8192 binary64:2048:65536 binary64:32768:1024 native64:32768:1024 binary64:65536:1024 binary64:65536:128 product:1024:65536 product:2048:65536 product:4096:16384 split64:2048:32768 split32:2048:65536 > gpurun_out/sweep_v5.txt 2>&1
cat gpurun_out/sweep_v5.txt
