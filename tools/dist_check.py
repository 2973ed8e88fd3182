"""Multi-GPU check (run under torchrun on a `gpurun --gpus N` box): the distributed four-step NTT and the
point-range-sharded MSM reproduce the single-GPU results bit for bit, and a chain of independent proofs
distributes one per GPU.  Prints one JSON line from rank 0."""
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

# ---- four-step NTT, k = 20 (BASELINE: all-to-all only at k >= 20) ----
log_n = int(os.environ.get("LOGN", "20"))
N = 1 << log_n
g = torch.Generator(device="cuda").manual_seed(1)
full = torch.randint(0, 1 << 62, (N, 4), dtype=torch.int64, device="cuda", generator=g)
full[:, 3] &= (1 << 59) - 1
dom = pkg.EvaluationDomain(2, log_n, ctx=ctx)
ref = full.clone()
ctx.fft_dev(ref, dom.omega, log_n, 1)
ctx.sync()
lo, hi = pkg.dist.shard_range(N, world, rank)
ops = pkg.dist.GpuNttOps(ctx)
out = pkg.dist.ntt_four_step(full[lo:hi].clone(), log_n, 0, ops)      # builds the twiddle cache
torch.cuda.synchronize()
res["ntt_ok"] = bool(torch.equal(out, ref[lo:hi]))
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = full[lo:hi].clone()
e0.record()
out = pkg.dist.ntt_four_step(x, log_n, 0, ops)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
res["ntt_four_step_ms"] = float(t.item())
e0.record(); ctx.fft_dev(ref, dom.omega, log_n, 1); e1.record(); torch.cuda.synchronize()
res["ntt_single_gpu_ms"] = e0.elapsed_time(e1)

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
allok = torch.tensor([int(res["ntt_ok"] and res["msm_ok"] and res["chain_ok"])], device="cuda")
dist.all_reduce(allok, op=dist.ReduceOp.MIN)
res["all_ranks_ok"] = bool(allok.item())
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
sys.exit(0 if res["all_ranks_ok"] else 1)
