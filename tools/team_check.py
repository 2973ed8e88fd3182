"""Team proving on real GPUs (run under torchrun on a `gpurun --gpus N` box): ONE create_proof spread over N GPUs —
MSM by point range, lagrange_to_coeff by column, coset transforms by residue class, h(X) by row block of the class-major
extended coset, NCCL collectives between.

  1. parity: for small circuits the team proof (with ZKC_TEAM_POISON=1: class blocks and rows a rank never computes or receives are 0xff) equals the
     single-GPU proof of the same context on every rank;
  2. timing: `WORKLOAD` (default agg_k20) single-GPU vs team, CUDA events, max over ranks, L2 flushed between proofs.
Prints one JSON line from rank 0; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

os.environ["ZKC_TEAM_POISON"] = "1"
pkg = graft.load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local)
ctx.use_torch_stream()
res = {"world": world, "parity": {}}

# ---- single-GPU proofs first (the context is not in a team yet) ----
cases = {
    "base_k8": pkg.workload.build(ctx, 8, 3, seed=5),
    "base_k12": pkg.workload.build(ctx, 12, 2, seed=6),
    "sha_bit_k10": pkg.workload.build(ctx, 10, 48, seed=4, shape="sha_bit"),
}
seed = pkg.seed_from_u64(99)
single = {name: [pkg.create_proof(w.pk, w.advice_dev, w.instances, seed, t, m) for t, m in (("blake2b", "shplonk"), ("keccak", "gwc"))]
          for name, w in cases.items()}
wl_name = os.environ.get("WORKLOAD", "agg_k20")
steps = int(os.environ.get("STEPS", "3"))
wl = bench.WORKLOADS[wl_name]
big = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100, shape=wl.get("shape", "base"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n):
    ms = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    return out, ms


big_seed = pkg.seed_from_u64(7)
prove_big = lambda: pkg.create_proof(big.pk, big.advice_dev, big.instances, big_seed)
prove_big()
big_single, ms_single = timed(prove_big, steps)
ctx.set_overlap(False)
_, ms_single_serial = timed(prove_big, 1)
ctx.set_overlap(True)

# ---- join the team ----
ctx.team_init()
res["team_info"] = list(ctx.team_info())
ok = True
for name, w in cases.items():
    got = [pkg.create_proof(w.pk, w.advice_dev, w.instances, seed, t, m) for t, m in (("blake2b", "shplonk"), ("keccak", "gwc"))]
    res["parity"][name] = bool(got == single[name])
    ok = ok and got == single[name]
prove_big()
big_team, ms_team = timed(prove_big, steps)
res["parity"][wl_name] = bool(big_team == big_single)
ok = ok and big_team == big_single
ctx.profile_enable(True)
ctx.profile_report()
prove_big()
prof = ctx.profile_report()
ctx.profile_enable(False)
allok = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(allok, op=dist.ReduceOp.MIN)
res["all_ranks_ok"] = bool(allok.item())
res["workload"] = wl_name
res["single_gpu_ms"] = min(ms_single)
res["single_gpu_one_stream_ms"] = min(ms_single_serial)
res["team_ms"] = min(ms_team)
res["team_ms_all"] = ms_team
res["speedup"] = res["single_gpu_ms"] / res["team_ms"]
res["team_phases_ms_rank0"] = {k: round(v["ms"], 3) for k, v in sorted(prof.items()) if not k.startswith("count:")}
if rank == 0:
    print(json.dumps(res))
ctx.team_leave()
dist.destroy_process_group()
sys.exit(0 if res["all_ranks_ok"] else 1)
