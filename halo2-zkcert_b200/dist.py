"""Multi-GPU sharding helpers (one process per GPU, torch.distributed).

The path shards without a data-path collective: whole proofs of a certificate chain are
independent (`/root/reference/README.md:28-32` runs one proof per CLI invocation), per-column
commitments are independent, and an MSM splits by point range exactly as the CPU `best_multiexp`
splits across threads.  The only exchange is the gather of results: proof bytes (a few KB) or one
partial sum per rank (96 B), done with all_gather on whatever backend the group uses (NCCL on the
GPUs, gloo in the CPU tests).

ONE proof spread over several GPUs (MSM by point range, transforms by column, h(X) by row block) is not
orchestrated from here: it lives inside the library (`csrc/dist.cu`, `zkc_team_*`, `Context.team_init`),
where the collectives run on the streams of the kernels they depend on.  The helpers below are the
proof-chain distributor / scheduler and the operator-level sharded MSM.  There is no NTT here: the size-2^extended_k coset
transforms are split by residue class inside the library (csrc/ntt.cu, dom_coeff_to_classes — the four-step decomposition
N1 = n, N2 = 2^(extended_k - k) of a zero-padded input, in which the exchange step disappears), natively and without a
transpose; the eager-torch four-step operator of round 1 was removed.
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


# ---- certificate-chain scheduler (BASELINE config 4) ----------------------------------------------------------------
# The chain's proofs are independent (2 x RSA k=17 + 2 x SHA256 k=19 for a 3-certificate chain:
# /root/reference/src/tests/x509_aggregation.rs:34-57, src/bin/cli.rs:385-390) but very unequal: one SHA proof costs about
# ten RSA proofs.  With more GPUs than proofs, or unequal proofs, "one proof per GPU" leaves most GPUs idle, so the GPUs are
# partitioned into TEAMS (zkc_team_*: one create_proof spread over the team) and the proofs are dealt to the teams.
def team_time(t1, g, serial_frac):
    """estimated time of a proof that takes t1 on one GPU when a team of g proves it (Amdahl with the measured
    non-scaling share: Fiat-Shamir-serial bucket phases, replicated scans, exchanges)"""
    return t1 * (serial_frac + (1.0 - serial_frac) / g)


def plan_chain(jobs, world):
    """jobs: [(name, t1_seconds, serial_frac)], world: number of GPUs.  Returns (teams, makespan): teams = list of
    (first_rank, size, [job indices in run order]); every rank belongs to exactly one team.  Exhaustive over the ways to cut
    `world` GPUs into at most len(jobs) contiguous teams (non-increasing sizes), longest-processing-time dealing of the jobs
    per cut; the search space is tiny (world <= 8, a handful of jobs)."""
    njobs = len(jobs)
    order = sorted(range(njobs), key=lambda i: -jobs[i][1])

    def cuts(total, parts, cap):
        if parts == 0:
            if total == 0:
                yield []
            return
        for first in range(min(cap, total - (parts - 1)), 0, -1):
            for rest in cuts(total - first, parts - 1, first):
                yield [first] + rest
    best = None
    for nteams in range(1, min(njobs, world) + 1):
        for sizes in cuts(world, nteams, world):
            load = [0.0] * nteams
            deal = [[] for _ in range(nteams)]
            for j in order:
                # the team that would finish this job first
                t = min(range(nteams), key=lambda q: (load[q] + team_time(jobs[j][1], sizes[q], jobs[j][2]), q))
                load[t] += team_time(jobs[j][1], sizes[t], jobs[j][2])
                deal[t].append(j)
            span = max(load)
            if best is None or span < best[0] - 1e-12:
                best = (span, sizes, deal)
    span, sizes, deal = best
    teams, first = [], 0
    for size, js in zip(sizes, deal):
        teams.append((first, size, js))
        first += size
    return teams, span


def team_of(teams, rank):
    for idx, (first, size, js) in enumerate(teams):
        if first <= rank < first + size:
            return idx
    raise ValueError("rank outside the plan")


def sum_partials(local_point, add_fn, group=None):
    """point-sharded MSM epilogue: gather one partial sum per rank and fold them with `add_fn`"""
    parts = gather_objects(local_point, group)
    acc = parts[0]
    for p in parts[1:]:
        acc = add_fn(acc, p)
    return acc


# ---- point-range-sharded MSM ----------------------------------------------------------
def msm_point_sharded(ctx, scalars_local, bases_local, n_local, group=None):
    """Each rank holds a contiguous slice of the bases (resident) and the matching scalar slice; local
    Pippenger on the GPU, all-gather of one 96-byte partial per rank, host-side sum.  Same result on
    every rank (normalised Jacobian, (1, 12))."""
    import numpy as np
    from . import api
    local = ctx.msm_dev(scalars_local, bases_local, n_local, 1) if n_local else np.zeros((1, 12), dtype=np.uint64)
    parts = gather_objects(local.tobytes(), group)
    pts = np.frombuffer(b"".join(parts), dtype=np.uint64).reshape(-1, 12)
    return api.g1_sum(pts)
