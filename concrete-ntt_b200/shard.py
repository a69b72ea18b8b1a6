"""Batch sharding across the GPUs of one box (one process per GPU, launched with torchrun).

Polynomials are independent and plans are pure functions of (n, p), so the batch is
split contiguously, plans are replicated per device, and the data path has NO collective.  `gather` is the
optional "give me everything on every rank" step (NCCL all-gather over NVLink on GPUs, gloo on CPU tensors);
it is never part of a timed NTT / polymul step.
"""
import torch
import torch.distributed as dist


def shard_range(batch, world_size, rank):
    """Contiguous split [lo, hi) of `batch` polynomials: rank r gets [r*B/G, (r+1)*B/G) (floor arithmetic), so
    shard sizes differ by at most one and concatenating the shards in rank order restores the batch."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (rank * batch) // world_size, ((rank + 1) * batch) // world_size


def shard(t, world_size=None, rank=None):
    """View of this rank's polynomials of a (batch, ...) tensor / array."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_range(t.shape[0], world_size, rank)
    return t[lo:hi]


def gather(local, batch, group=None):
    """All-gather the per-rank results (shapes (hi-lo, ...)) into the full (batch, ...) tensor on every rank.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
