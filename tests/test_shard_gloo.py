"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo groups.  The data path has no collective -- each rank
transforms its own contiguous shard -- so what is tested is the partition and the optional result gather.
The per-shard "compute" is the CPU oracle (the product has no CPU path); on the GPU box the same helpers are
driven by bench.py under torchrun."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cntt = importlib.import_module("concrete-ntt_b200")
        from oracle import oracle as O
        n, p = 64, 1062862849
        g = np.random.Generator(np.random.PCG64(123))
        full = g.integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)   # same on every rank
        lo, hi = cntt.shard.shard_range(batch, world, rank)
        mine = cntt.shard.shard(torch.from_numpy(full.view(np.int32)))
        assert mine.shape[0] == hi - lo
        plan = O.Plan32.try_new(n, p)
        local = mine.numpy().view(np.uint32).copy()
        if local.shape[0]:
            plan.fwd(local)
        gathered = cntt.shard.gather(torch.from_numpy(local.view(np.int32)), batch)
        ref = plan.fwd(full.copy())
        ok = bool((gathered.numpy().view(np.uint32) == ref).all())
        q.put((rank, lo, hi, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 10), (2, 7), (3, 8), (2, 1)])
def test_shard_and_gather_gloo(world, batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    # shards tile [0, batch) in rank order and every rank reconstructed the full result
    assert res[0][1] == 0 and res[-1][2] == batch
    for a, b in zip(res, res[1:]):
        assert a[2] == b[1]
    assert all(r[3] for r in res)


def test_shard_range_properties():
    import importlib
    sys.path.insert(0, ROOT)
    cntt = importlib.import_module("concrete-ntt_b200")
    for batch in (0, 1, 5, 65536, 1024, 1023):
        for world in (1, 2, 4, 8):
            spans = [cntt.shard.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
