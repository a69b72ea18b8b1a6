#!/bin/bash
# r02 session F (8 GPUs): PCIe / host-memory probe on all ranks at once, then the bench at 8 ranks
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py > gpurun_out/r02_pcie_probe_${N}gpu.txt 2>&1
cat gpurun_out/r02_pcie_probe_${N}gpu.txt | grep -v Warning | tail -5
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/r02_host_${N}gpu.txt; nvidia-smi topo -m >> gpurun_out/r02_host_${N}gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${N}gpu.json") if x.startswith("{")]
d=json.loads(l[-1]); print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["e2e"]["value"], d["e2e_fused"]["value"], d["numa_node_rank0"])
PY
