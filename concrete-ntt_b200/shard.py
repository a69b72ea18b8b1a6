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


class PeerGather:
    """Single-device output without a collective after the kernel: the result rows of every rank are written by the
    kernels THEMSELVES into the root rank's HBM over NVLink peer memory.

    Every rank allocates the same symmetric (batch, n) buffer (torch symmetric memory: CUDA VMM allocations whose
    handles are exchanged once at construction); `dest()` is this rank's shard [lo, hi) of the ROOT's buffer, mapped
    into this process -- a plain device pointer, so it can be handed to any out-of-place entry point of the C ABI
    (`negacyclic_polymul(prod=gather.dest(), ...)`, `cntt_native_polymul`) or be the target of a device copy.  The
    stores of the kernel then travel through NVSwitch while its butterflies run; `wait()` is a stream-ordered
    barrier over the group after which `result()` on the root holds the whole batch.  (the
    gather is optional and never part of a timed NTT step; `gather()` above is the NCCL form of the same thing.)
    """

    def __init__(self, batch, n, dtype, root=0, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank, self.root = dist.get_world_size(self.group), dist.get_rank(self.group), root
        self.batch, self.n, self.dtype = batch, n, dtype
        self.local = symm.empty((batch, n), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.local, self.group)
        self.lo, self.hi = shard_range(batch, self.world, self.rank)
        self._dest = self.hdl.get_buffer(root, (self.hi - self.lo, n), dtype, self.lo * n)

    def dest(self):
        """Rows [lo, hi) of the root's buffer (peer memory unless this rank is the root)."""
        return self._dest

    def wait(self):
        """Barrier on the current stream: returns once every rank's preceding work on its stream is visible."""
        self.hdl.barrier()

    def result(self):
        """The (batch, n) buffer of this rank; complete on the root after wait()."""
        return self.local
