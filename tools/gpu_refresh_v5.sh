mkdir -p gpurun_out
run() { local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none -k regex:"$rx" -s 1 -c 2 -f -o gpurun_out/prof_$name python tools/prof_driver.py "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_sum_$name.txt 2>&1; rm -f gpurun_out/prof_$name.ncu-rep gpurun_out/ncu_$name.log; }
run polymul64_2048  'k_polymul_fused'  polymul64 32768 2048
run polymul32_2048  'k_polymul_fused'  polymul32 32768 2048
run polymulb64_2048 'k_polymul_fused'  polymulb64 32768 2048
bash tools/gpu_stalls.sh pm64 k_polymul_fused polymul64 32768 2048 > /dev/null 2>&1; rm -f gpurun_out/st_pm64.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; cut -c1-200 gpurun_out/bench_v5.json
timeout 600 python tools/time_polymul.py p32:256:262144 p32:1024:65536 p32:2048:65536 p32:4096:32768 p32:16384:16384 p32:65536:4096 p64s:1024:65536 p64s:2048:65536 p64:2048:65536 p64s:4096:32768 p64s:16384:8192 p64s:65536:2048 native64:1024:65536 native64:2048:65536 native64:4096:16384 native32:2048:65536 native128:2048:16384 native128:4096:8192 binary64:2048:65536 binary64:32768:1024 native64:32768:1024 binary64:65536:1024 binary64:65536:128 product:1024:65536 product:2048:65536 product:4096:16384 split64:2048:32768 split32:2048:65536 > gpurun_out/sweep_v5.txt 2>&1
grep -E "native64 n=2048|p64s n=2048" gpurun_out/sweep_v5.txt
