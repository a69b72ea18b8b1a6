#!/bin/bash
# r02, two GPUs: the tests a one-GPU box skips (peer gather over NVLink, host_multi across devices), the bench at N=2, and the
# pre-transformed binary polymul figures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "peer or multi or shard" > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_2gpu.log; tail -4 gpurun_out/r02_pytest_2gpu.log
timeout 600 python tools/time_polymul.py preb64:2048:65536 preb32:2048:65536 pre64:2048:65536 binary64:2048:65536 2>&1 | tee gpurun_out/r02_preb.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/r02_bench_2gpu.json
