"""Multi-GPU sharding helpers (one process per GPU, torch.distributed).

The path shards without a data-path collective: whole proofs of a certificate chain are
independent (`/root/reference/README.md:28-32` runs one proof per CLI invocation), per-column
commitments are independent, and an MSM splits by point range exactly as the CPU `best_multiexp`
splits across threads.  The only exchange is the gather of results: proof bytes (a few KB) or one
partial sum per rank (96 B), done with all_gather on whatever backend the group uses (NCCL on the
GPUs, gloo in the CPU tests).
"""
import torch.distributed as dist


def shard_range(total, world, rank):
    """contiguous [start, end) of `total` items for `rank` — the point-range split of best_multiexp"""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def assign_round_robin(n_items, world, rank):
    """indices of the independent jobs (proofs, columns) this rank owns"""
    return list(range(rank, n_items, world))


def gather_objects(local, group=None):
    """every rank receives [rank0's object, rank1's object, ...]"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def prove_chain(jobs, prove_fn, group=None):
    """Distribute independent proving jobs one-per-GPU (BASELINE config 4: 2 RSA + 2 SHA256 proofs of a
    3-certificate chain).  `prove_fn(job) -> bytes` runs on this rank's GPU.  Returns the proofs in job
    order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = assign_round_robin(len(jobs), world, rank)
    local = {i: prove_fn(jobs[i]) for i in mine}
    merged = {}
    for part in gather_objects(local, group):
        merged.update(part)
    assert sorted(merged) == list(range(len(jobs)))
    return [merged[i] for i in range(len(jobs))]


def sum_partials(local_point, add_fn, group=None):
    """point-sharded MSM epilogue: gather one partial sum per rank and fold them with `add_fn`"""
    parts = gather_objects(local_point, group)
    acc = parts[0]
    for p in parts[1:]:
        acc = add_fn(acc, p)
    return acc
