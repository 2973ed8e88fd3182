"""Multi-GPU sharding helpers (one process per GPU, torch.distributed).

The path shards without a data-path collective: whole proofs of a certificate chain are
independent (`/root/reference/README.md:28-32` runs one proof per CLI invocation), per-column
commitments are independent, and an MSM splits by point range exactly as the CPU `best_multiexp`
splits across threads.  The only exchange is the gather of results: proof bytes (a few KB) or one
partial sum per rank (96 B), done with all_gather on whatever backend the group uses (NCCL on the
GPUs, gloo in the CPU tests).

ONE proof spread over several GPUs (MSM by point range, transforms by column, h(X) by row block) is not
orchestrated from here: it lives inside the library (`csrc/dist.cu`, `zkc_team_*`, `Context.team_init`),
where the collectives run on the streams of the kernels they depend on.  The helpers below remain the
operator-level building blocks (sharded MSM, four-step NTT) and the proof-chain distributor.
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


# ---- four-step NTT with all-to-all transposes (k >= 20) --------------------------------
class GpuNttOps:
    """local pieces of the distributed NTT on one GPU, built from the single-GPU C ABI"""

    def __init__(self, ctx):
        self.ctx = ctx
        self._tw = {}

    def fft_rows(self, t, log_len, inverse):
        """t: [rows, 2^log_len, 4] int64 CUDA, contiguous; in-place NTT of every row"""
        rows = t.shape[0]
        dom = self._dom(log_len)
        omega = dom.omega_inv if inverse else dom.omega
        self.ctx.fft_dev(t.view(-1, 4), omega, log_len, rows)
        return t

    def _dom(self, log_len):
        from . import api
        key = ("dom", log_len)
        if key not in self._tw:
            self._tw[key] = api.EvaluationDomain(2, log_len, ctx=self.ctx)
        return self._tw[key]

    def twiddles(self, row0, nrows, length, log_n, inverse):
        """T[r][c] = w_N^((row0 + r) * c), w_N the canonical 2^log_n-th root (or its inverse); cached"""
        import torch
        key = ("tw", row0, nrows, length, log_n, inverse)
        if key not in self._tw:
            dom = self._dom(log_n)
            w = dom.omega_inv if inverse else dom.omega
            from .workload import to_mont_dev, to_host
            one = to_host(to_mont_dev(self.ctx, [1]))
            dev = "cuda:%d" % self.ctx.device
            bases = torch.empty((row0 + nrows, 4), dtype=torch.int64, device=dev)
            self.ctx.powers_dev(bases, w, one)            # w^i, i < row0 + nrows
            bases_h = to_host(bases)
            T = torch.empty((nrows, length, 4), dtype=torch.int64, device=dev)
            for r in range(nrows):
                self.ctx.powers_dev(T[r], bases_h[row0 + r:row0 + r + 1], one)
            self.ctx.sync()
            self._tw[key] = T
        return self._tw[key]

    def mul(self, a, b):
        import torch
        out = torch.empty_like(a)
        self.ctx.field_vec_op_dev("fr", "mul", a.reshape(-1, 4), b.reshape(-1, 4), out.view(-1, 4))
        return out

    def scale(self, a, s_mont):
        import torch
        b = torch.from_numpy(s_mont.view("int64")).to(a.device).expand(a.numel() // 4, 4).contiguous()
        return self.mul(a, b.view(a.shape))


def _all_to_all(t, group):
    import torch
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    out = torch.empty_like(t)
    dist.all_to_all_single(out, t.contiguous(), group=group)
    return out


def ntt_four_step(x_local, log_n, inverse, ops, group=None):
    """Distributed size-2^log_n NTT over G ranks; natural order in and out, rank g holding the contiguous
    block [g N/G, (g+1) N/G) (x_local: [N/G, 4] int64 on the ops' device).

    N = N1 * N2 with n = n1*N2 + n2 and k = k1 + N1*k2:
      transpose (all-to-all)  -> rank holds a block of columns n2, all n1
      local NTT over n1 (length N1) for each owned column, times w_N^(n2*k1)
      transpose (all-to-all)  -> rank holds a block of k1, all n2
      local NTT over n2 (length N2) for each owned k1
      transpose (all-to-all)  -> natural order: rank holds a block of k2, all k1
    The inverse transform applies the same steps with w^-1; scaling by 1/N is the caller's."""
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    g = dist.get_rank(group) if dist.is_initialized() else 0
    N = 1 << log_n
    l1 = (log_n + 1) // 2
    l2 = log_n - l1
    N1, N2 = 1 << l1, 1 << l2
    assert N1 % G == 0 and N2 % G == 0 and x_local.shape[0] == N // G
    r1, c2 = N1 // G, N2 // G
    # local rows n1 in block g: [r1][N2] -> send column block j to rank j
    a = x_local.view(r1, G, c2, 4).permute(1, 0, 2, 3).contiguous()          # [G][r1][c2]
    a = _all_to_all(a.view(G * r1 * c2, 4), group).view(G * r1, c2, 4)       # [N1][c2]: all n1, my columns
    cols = a.permute(1, 0, 2).contiguous()                                   # [c2][N1]
    cols = ops.fft_rows(cols, l1, inverse)                                   # Y[n2][k1]
    cols = ops.mul(cols, ops.twiddles(g * c2, c2, N1, log_n, inverse))
    # now distribute by k1: [c2][G][r1] -> rank j gets k1 block j
    b = cols.view(c2, G, r1, 4).permute(1, 0, 2, 3).contiguous()             # [G][c2][r1]
    b = _all_to_all(b.view(G * c2 * r1, 4), group).view(G * c2, r1, 4)       # [N2][r1]: all n2, my k1
    rows = b.permute(1, 0, 2).contiguous()                                   # [r1][N2]
    rows = ops.fft_rows(rows, l2, inverse)                                   # X[k1][k2], k = k1 + N1*k2
    # natural order: rank j owns k2 block j, all k1 -> [G][r1][c2] -> all-to-all -> [N1][c2] -> [c2][N1]
    c = rows.view(r1, G, c2, 4).permute(1, 0, 2, 3).contiguous()
    c = _all_to_all(c.view(G * r1 * c2, 4), group).view(G * r1, c2, 4)       # [k1 all][my k2]
    return c.permute(1, 0, 2).contiguous().view(N // G, 4)                   # index = k2_local*N1 + k1
