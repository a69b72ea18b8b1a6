"""Two ranks, one GPU each: the native64 polymul kernels store their product rows straight into rank 0's HBM through
NVLink peer memory (shard.PeerGather); rank 0's buffer must equal the CPU oracle's product of the whole batch.
Needs two GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, batch, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cntt = importlib.import_module("concrete-ntt_b200")
        g = np.random.Generator(np.random.PCG64(99))
        lhs = g.integers(0, 2**64, size=(batch, n), dtype=np.uint64)     # same on every rank
        rhs = g.integers(0, 2**64, size=(batch, n), dtype=np.uint64)
        lo, hi = cntt.shard.shard_range(batch, world, rank)
        plan = cntt.native64.Plan32.try_new(n, device=rank)
        dl = torch.from_numpy(lhs[lo:hi].view(np.int64).copy()).cuda()
        dr = torch.from_numpy(rhs[lo:hi].view(np.int64).copy()).cuda()
        pg = cntt.shard.PeerGather(batch, n, torch.int64, root=0)
        pg.result().zero_()
        torch.cuda.synchronize(); dist.barrier()
        plan.negacyclic_polymul(pg.dest(), dl, dr)
        pg.wait()
        torch.cuda.synchronize()
        if rank == 0:
            q.put(("result", pg.result().cpu().numpy().view(np.uint64), lhs, rhs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(256, 7), (2048, 64)])
def test_polymul_stores_into_root_over_peer_memory(oracle, torch_cuda, n, batch):
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    tag, got, lhs, rhs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = oracle.Native.try_new(n, 64, binary=False).negacyclic_polymul(lhs, rhs)
    assert (got == ref).all()
