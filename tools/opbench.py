"""Kernel micro-shapes of BASELINE.md §4 on one B200: MSM points/s for N in {2^15, 2^17, 2^19, 2^22} x {uniform, bits,
<2^16} scalars against the resident SRS, and NTT batches 16x2^17, 115x2^19, 8x2^22 plus the extended-domain sizes,
each with the fraction of the measured integer roofline (Fr products/s) and of the measured HBM copy bandwidth.
Writes gpurun_out/opbench.json (copied to profiles/ by hand)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
ctx = pkg.Context(0)
ctx.use_torch_stream()
FE_MUL_PEAK = json.load(open(os.path.join(ROOT, "profiles", "r01_ffbench.json")))["fr_mul_per_s"]
IMADW_PEAK = json.load(open(os.path.join(ROOT, "profiles", "r01_ffbench.json")))["imad_wide_per_s"]
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def rand_fr(n, seed, kind="uniform"):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 59) - 1
    if kind == "uniform":
        return t
    t[:, 1:] = 0
    t[:, 0] &= 1 if kind == "bits" else 0xFFFF
    out = torch.empty_like(t)
    ctx.field_vec_op_dev("fr", "from_canonical", t, None, out)
    return out


def timeit(fn, reps=5, warm=3):
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


res = {"peaks": {"fr_mul_per_s": FE_MUL_PEAK, "imad_wide_per_s": IMADW_PEAK, "hbm_gbs": HBM}, "msm": {}, "ntt": {}}
s = pkg.api.fr_random_stream(bytes(32), 1)
for k in [int(x) for x in os.environ.get("KS", "15,17,19,22").split(",") if x]:
    n = 1 << k
    params = pkg.ParamsKZG.setup(k, s, ctx=ctx)
    for kind in ("uniform", "bits", "u16"):
        for ncols in ((1, 16) if k <= 17 else (1, 4) if k <= 19 else (1,)):
            sc = rand_fr(n * ncols, 1, kind)
            ms = timeit(lambda: params.commit_dev(sc, n, ncols, 0))
            ctx.profile_enable(True); ctx.profile_report()
            params.commit_dev(sc, n, ncols, 0)
            prof = ctx.profile_report(); ctx.profile_enable(False)
            madds = prof.get("count:msm.madds", {"n": 0})["n"]
            acc = prof.get("msm.accum", {"ms": 0})["ms"]
            res["msm"]["k%d_%s_x%d" % (k, kind, ncols)] = {
                "ms": round(ms, 4), "points_per_s": n * ncols / ms * 1e3, "madds": madds, "accum_ms": round(acc, 4),
                "accum_frac_of_imadw_peak": (madds * 1280 / (acc * 1e-3) / IMADW_PEAK) if acc else None,
                "phases_ms": {a: round(b["ms"], 4) for a, b in prof.items() if a.startswith("msm.")}}
            del sc
    del params
    torch.cuda.empty_cache()
shapes = [(17, 16), (19, 115), (22, 8), (15, 30)]
for k, ncols in shapes:
    n = 1 << k
    dom = pkg.EvaluationDomain(4, k, ctx=ctx)
    a = rand_fr(n * ncols, 2)
    ms = timeit(lambda: dom.lagrange_to_coeff_dev(a, ncols))
    muls = ncols * (n / 2 * k + n + (n if k > 11 else 0) * (1 if k <= 18 else 2))
    res["ntt"]["intt_%dx2^%d" % (ncols, k)] = {"ms": round(ms, 4), "GBps_algorithmic": 64.0 * n * ncols / ms / 1e6, "frac_hbm": 64.0 * n * ncols / ms / 1e6 / HBM,
                                              "fr_mul_per_s": muls / ms * 1e3, "frac_int": muls / ms * 1e3 / FE_MUL_PEAK}
    ek, en = dom.extended_k, dom.extended_n
    ce = min(ncols, 16 if k <= 19 else 4)
    ext = torch.empty((en * ce, 4), dtype=torch.int64, device="cuda")
    ms = timeit(lambda: dom.coeff_to_extended_dev(a[: n * ce], ext, ce))
    muls = ce * (en / 2 * ek + n + en * (1 if ek <= 18 else 2))
    res["ntt"]["coset_%dx2^%d_to_2^%d" % (ce, k, ek)] = {"ms": round(ms, 4), "GBps_algorithmic": 32.0 * (n + en) * ce / ms / 1e6,
                                                        "frac_hbm": 32.0 * (n + en) * ce / ms / 1e6 / HBM, "fr_mul_per_s": muls / ms * 1e3,
                                                        "frac_int": muls / ms * 1e3 / FE_MUL_PEAK}
    ms = timeit(lambda: dom.extended_to_coeff_dev(ext, ce))
    muls = ce * (en / 2 * ek + en + en * (1 if ek <= 18 else 2))
    res["ntt"]["ext_intt_%dx2^%d" % (ce, ek)] = {"ms": round(ms, 4), "GBps_algorithmic": 64.0 * en * ce / ms / 1e6, "frac_hbm": 64.0 * en * ce / ms / 1e6 / HBM,
                                                "fr_mul_per_s": muls / ms * 1e3, "frac_int": muls / ms * 1e3 / FE_MUL_PEAK}
    del a, ext, dom
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "opbench.json"), "w"), indent=1)
