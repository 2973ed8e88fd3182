"""Operator micro-shapes of BASELINE.md §4 on one B200: MSM pts/s and NTT throughput with per-kernel
CUDA-event breakdown (zkc_profile_*).  Development tool; bench.py is the judged harness."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
ctx = pkg.Context(0)
ctx.use_torch_stream()


def rand_fr(n, seed, kind="uniform"):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 59) - 1
    if kind == "bits":
        t[:, 1:] = 0
        t[:, 0] &= 1
        out = torch.empty_like(t)
        ctx.field_vec_op_dev("fr", "from_canonical", t, None, out)
        return out
    return t


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
s = np.array([[0x1234567, 0, 0, 0]], dtype=np.uint64)
for k in [int(x) for x in os.environ.get("KS", "15,17,19").split(",")]:
    n = 1 << k
    t0 = time.time()
    params = pkg.ParamsKZG.setup(k, s, ctx=ctx)
    ctx.sync()
    res["srs_setup_k%d_s" % k] = time.time() - t0
    for kind in ("uniform", "bits"):
        for ncols in (1, 16):
            if k >= 19 and ncols > 4:
                continue
            sc = rand_fr(n * ncols, 1, kind)
            ms = timeit(lambda: params.commit_dev(sc, n, ncols, 0))
            ctx.profile_enable(True)
            params.commit_dev(sc, n, ncols, 0)
            prof = ctx.profile_report()
            ctx.profile_enable(False)
            res["commit_k%d_%s_x%d" % (k, kind, ncols)] = {"ms": ms, "pts_per_s": n * ncols / ms * 1e3,
                                                         "prof": {a: round(b["ms"], 4) for a, b in prof.items()}}
    del params
    for j, ncols in ((4, 1), (4, 16)):
        if k >= 19 and ncols > 4:
            continue
        dom = pkg.EvaluationDomain(j, k, ctx=ctx)
        a = rand_fr(n * ncols, 2)
        ms = timeit(lambda: dom.lagrange_to_coeff_dev(a, ncols))
        res["intt_k%d_x%d" % (k, ncols)] = {"ms": ms, "GBps": 64.0 * n * ncols / ms / 1e6}
        ext = torch.empty((dom.extended_n * ncols, 4), dtype=torch.int64, device="cuda")
        ms = timeit(lambda: dom.coeff_to_extended_dev(a, ext, ncols))
        res["coset_ntt_k%d_to_%d_x%d" % (k, dom.extended_k, ncols)] = {"ms": ms, "GBps": 32.0 * (n + dom.extended_n) * ncols / ms / 1e6}
        ms = timeit(lambda: dom.extended_to_coeff_dev(ext, ncols))
        res["ext_intt_%d_x%d" % (dom.extended_k, ncols)] = {"ms": ms, "GBps": 64.0 * dom.extended_n * ncols / ms / 1e6}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "opbench.json"), "w"), indent=1)
