"""Single-device output of a sharded native64 polymul: three ways, timed on the device (max over ranks).

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/gather_probe.py [n] [batch]

  local     every rank writes its product into its own HBM (the sharded hot path, no exchange)
  peer      every rank's polymul kernel stores straight into the ROOT's HBM over NVLink (shard.PeerGather):
            compute and gather are one launch
  nccl      local polymul, then shard.gather (NCCL all-gather) -- the collective form
The peer rows are compared with the locally computed product (bit-exact), the NCCL result with both.
One JSON line from rank 0.
"""
import importlib, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cntt = importlib.import_module("concrete-ntt_b200")


def timed(fn, reps, pre=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    plan = cntt.native64.Plan32.try_new(n, device=lrank)
    lo, hi = cntt.shard.shard_range(batch, world, rank)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    big = 2**63 - 1
    lhs = torch.randint(-big - 1, big, (hi - lo, n), dtype=torch.int64, device="cuda", generator=g)
    rhs = torch.randint(-big - 1, big, (hi - lo, n), dtype=torch.int64, device="cuda", generator=g)
    local = torch.empty_like(lhs)
    reps = 10

    t_local = timed(lambda: plan.negacyclic_polymul(local, lhs, rhs), reps)

    pg = cntt.shard.PeerGather(batch, n, torch.int64, root=0)
    dest = pg.dest()

    def peer_step():
        plan.negacyclic_polymul(dest, lhs, rhs)
        pg.wait()
    t_peer = timed(peer_step, reps)
    ok_peer = bool(torch.equal(dest, local))            # read back through the peer mapping
    full = None

    def nccl_step():
        nonlocal full
        plan.negacyclic_polymul(local, lhs, rhs)
        full = cntt.shard.gather(local, batch)
    t_nccl = timed(nccl_step, reps)
    ok_nccl = bool(torch.equal(full[lo:hi], local))
    if rank == 0:
        ok_nccl = ok_nccl and bool(torch.equal(full, pg.result()))
    flags = torch.tensor([int(ok_peer), int(ok_nccl)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    # plain device copy of the finished shard into the root's buffer (what a caller would do without the fused store)
    t_copy = timed(lambda: (dest.copy_(local), pg.wait()), reps)

    if rank == 0:
        shard_bytes = (hi - lo) * n * 8
        print(json.dumps({
            "workload": "native64 negacyclic polymul n=%d, batch %d split over %d GPUs, output wanted on GPU 0" % (n, batch, world),
            "ms_local_no_gather": t_local, "ms_peer_store_fused": t_peer, "ms_local_then_nccl_all_gather": t_nccl,
            "ms_copy_shard_to_root_only": t_copy,
            "polymuls_per_s_local": batch / t_local * 1e3, "polymuls_per_s_peer": batch / t_peer * 1e3,
            "polymuls_per_s_nccl": batch / t_nccl * 1e3,
            "shard_bytes": shard_bytes, "peer_rows_bit_exact": bool(flags[0].item()), "nccl_rows_bit_exact": bool(flags[1].item()),
            "n_gpus": world}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
