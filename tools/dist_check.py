"""Multi-GPU check (run under torchrun on a `gpurun --gpus N` box): the point-range-sharded MSM operator reproduces the
single-GPU result bit for bit, and a chain of independent proofs distributes one per GPU (team proving: tools/team_check.py).  Prints one JSON line from rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local)
ctx.use_torch_stream()
res = {"world": world}

g = torch.Generator(device="cuda").manual_seed(1)
# ---- point-sharded MSM ----
k = 16
n = 1 << k
params = pkg.ParamsKZG.setup(k, pkg.api.fr_random_stream(bytes(32), 1), ctx=ctx)
bases = torch.from_numpy(params.get_g(0).view(np.int64)).cuda()
sc = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 59) - 1
whole = ctx.msm_dev(sc, bases, n, 1)
lo, hi = pkg.dist.shard_range(n, world, rank)
part = pkg.dist.msm_point_sharded(ctx, sc[lo:hi].contiguous(), bases[lo:hi].contiguous(), hi - lo)
res["msm_ok"] = bool(np.array_equal(part, whole))

# ---- chain of independent proofs, one per GPU ----
w = pkg.workload.build(ctx, 10, 2, seed=7)           # same circuit on every rank; jobs differ by RNG seed
jobs = [pkg.seed_from_u64(100 + i) for i in range(2 * world)]
t0 = time.perf_counter()
proofs = pkg.dist.prove_chain(jobs, lambda seed: pkg.create_proof(w.pk, w.advice_dev, w.instances, seed))
res["chain_s"] = time.perf_counter() - t0
mine = pkg.create_proof(w.pk, w.advice_dev, w.instances, jobs[0])
res["chain_ok"] = bool(len(proofs) == len(jobs) and proofs[0] == mine and all(len(p) == len(mine) for p in proofs))
allok = torch.tensor([int(res["msm_ok"] and res["chain_ok"])], device="cuda")
dist.all_reduce(allok, op=dist.ReduceOp.MIN)
res["all_ranks_ok"] = bool(allok.item())
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
sys.exit(0 if res["all_ranks_ok"] else 1)
