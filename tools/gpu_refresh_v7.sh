#!/bin/bash
# r01 v7 evidence pass (after the 256-bit global accesses): ncu summaries of the kernels whose loads / stores changed,
# launch list of the bench command, full parity suite, smoke, both bench arms, sweep.  Outputs: gpurun_out/.
mkdir -p gpurun_out
run() { local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none -k regex:"$rx" -s 1 -c 2 -f -o gpurun_out/prof_$name python tools/prof_driver.py "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_sum_$name.txt 2>&1; rm -f gpurun_out/prof_$name.ncu-rep gpurun_out/ncu_$name.log; }
run polymul128_4096 'k_polymul_fused' polymul128 2048 4096
run ntt32_8192      'k_ntt_cta'       ntt32 8192 8192
run ntt32_1024      'k_ntt_cta'       ntt32 65536 1024
run ntt32_4096      'k_ntt_cta'       ntt32 16384 4096
run ntt64s_2048     'k_ntt_cta'       ntt64 65536 2048
run split64_2048    'k_native_fwd_fused' split64 32768 2048
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_v7.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench_v7.csv > gpurun_out/launches_bench_v7_summary.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_v7_reference_arm.json 2> gpurun_out/bench_v7_ref.err; cut -c1-200 gpurun_out/bench_v7_reference_arm.json
timeout 600 python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; cat gpurun_out/bench_v7.json
timeout 600 python tools/time_polymul.py p32:256:262144 p32:1024:65536 p32:2048:65536 p32:4096:32768 p32:8192:32768 p32:16384:16384 p32:65536:4096 p64s:1024:65536 p64s:2048:65536 p64:2048:65536 p64s:4096:32768 p64s:8192:16384 p64s:16384:8192 p64s:65536:2048 native64:1024:65536 native64:2048:65536 native64:4096:16384 native32:2048:65536 native128:1024:32768 native128:2048:16384 native128:4096:8192 binary64:2048:65536 binary128:4096:8192 binary64:32768:1024 native64:32768:1024 binary64:65536:1024 binary64:65536:128 product:1024:65536 product:2048:65536 product:4096:16384 split64:2048:32768 split32:2048:65536 > gpurun_out/sweep_v7.txt 2>&1
cat gpurun_out/sweep_v7.txt
