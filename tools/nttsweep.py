"""NTT pass-count sweep on one B200: batches of the BASELINE shapes with two passes allowed up to 2^18 (three above) vs up
to 2^22.  Prints ms per shape and setting; exact results are compared between the settings."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
ctx = pkg.Context(0)
ctx.use_torch_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def rand_fr(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 59) - 1
    return t


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


res = {}
for k, ncols in [(17, 16), (19, 32), (20, 8), (22, 4)]:
    n = 1 << k
    dom = pkg.EvaluationDomain(4, k, ctx=ctx)
    en = dom.extended_n
    ce = min(ncols, 16 if k <= 19 else 4)
    a0 = rand_fr(n * ncols, 2)
    outs = {}
    for setting in ("18", "20", "22"):
        ctx.set_tunable("ntt_two_pass_max", int(setting))
        a = a0.clone()
        ext = torch.empty((en * ce, 4), dtype=torch.int64, device="cuda")
        r = {}
        r["intt_%dx2^%d" % (ncols, k)] = round(timeit(lambda: dom.lagrange_to_coeff_dev(a, ncols)), 4)
        r["coset_%dx2^%d" % (ce, dom.extended_k)] = round(timeit(lambda: dom.coeff_to_extended_dev(a[: n * ce], ext, ce)), 4)
        r["ext_intt_%dx2^%d" % (ce, dom.extended_k)] = round(timeit(lambda: dom.extended_to_coeff_dev(ext, ce)), 4)
        # exactness across settings: one fresh transform chain
        b = a0[: n * ce].clone()
        dom.lagrange_to_coeff_dev(b, ce)
        dom.coeff_to_extended_dev(b, ext, ce)
        dom.extended_to_coeff_dev(ext, 1)
        torch.cuda.synchronize()
        outs[setting] = (b.clone(), ext[:en].clone())
        res.setdefault("k%d" % k, {})[setting] = r
        del a, ext
    for sname in ("20", "22"):
        assert torch.equal(outs["18"][0], outs[sname][0]) and torch.equal(outs["18"][1], outs[sname][1]), "pass-count settings disagree"
    del a0, outs
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "nttsweep.json"), "w"), indent=1)
