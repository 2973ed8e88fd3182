import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as graft
pkg = graft.load_package()
"""Development helper: create_proof time and MSM phase times for several window widths (msm_c_pre tunable).
    WL=rsa_k17 CS=0,16,17 TS=0,16 OCCS=0,3 python tools/sweep_c.py   (window width, entries per accumulate thread, accumulate CTAs/SM)"""
import bench
res = {}
wl = bench.WORKLOADS[os.environ.get("WL", "rsa_k17")]
import itertools
for c, T, occ in itertools.product([int(x) for x in os.environ.get("CS", "0,16,17").split(",")], [int(x) for x in os.environ.get("TS", "0").split(",")],
                                   [int(x) for x in os.environ.get("OCCS", "0").split(",")]):
    ctx = pkg.Context(0)
    ctx.use_torch_stream()
    ctx.set_tunable("msm_c_pre", c)
    ctx.set_tunable("msm_T", T)
    ctx.set_tunable("msm_accum_occ", occ)
    w = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100, shape=wl.get("shape", "base"))
    seeds = [pkg.seed_from_u64(i) for i in range(8)]
    for s in seeds[:3]:
        pkg.create_proof(w.pk, w.advice_dev, w.instances, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in seeds[3:]:
        p = pkg.create_proof(w.pk, w.advice_dev, w.instances, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ctx.set_overlap(False); ctx.profile_enable(True); ctx.profile_report()
    for s in seeds[3:]:
        pkg.create_proof(w.pk, w.advice_dev, w.instances, s)
    prof = ctx.profile_report(); ctx.profile_enable(False)
    c = "c%d_T%d_occ%d" % (c, T, occ)
    res[c] = {"ms": round(ms, 3), **{k: round(v["ms"] / 5, 3) for k, v in prof.items() if k.startswith("msm.")}}
    print(c, json.dumps(res[c]), flush=True)
    del w
    torch.cuda.empty_cache()
