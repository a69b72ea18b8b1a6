"""PCIe / host-memory ceiling of the whole box: every rank (one per GPU, torchrun) moves pinned 1 GiB buffers H2D, D2H and both ways
AT THE SAME TIME as all other ranks; reports per-rank and aggregate GB/s, with and without binding each rank to the NUMA node of
its GPU before the pinned buffers are allocated (first touch).  usage: torchrun --nproc-per-node N tools/pcie_probe_multi.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def probe(tag):
    n = 1 << 30
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h1.fill_(1); h2.fill_(2)          # first touch on this rank's cores
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def up():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def both():
        up(); down()

    out = []
    for name, fn, vol in (("H2D", up, n), ("D2H", down, n), ("both", both, 2 * n)):
        fn(); barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        gbs = vol / dt / 1e9
        t = torch.tensor([gbs, gbs], device="cuda", dtype=torch.float64)
        if world > 1:
            s = t.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
            m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MIN)
            out.append((name, float(s[0]), float(m[0])))
        else:
            out.append((name, gbs, gbs))
        barrier()
    if rank == 0:
        print("%s, %d ranks at once: " % (tag, world) + "   ".join("%s %.1f GB/s aggregate (slowest rank %.1f)" % o for o in out), flush=True)
    del h1, h2, d1, d2


probe("unbound (default placement)")
node = bind_to_gpu_numa_node(local)
if rank == 0:
    print("rank 0 bound to NUMA node %s" % (node,), flush=True)
probe("bound to the GPU's NUMA node")
if world > 1:
    dist.destroy_process_group()
