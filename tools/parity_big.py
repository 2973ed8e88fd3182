"""Byte parity at the BASELINE sizes the test-suite cannot afford: the device prover's proof vs the CPU oracle's for the
zkEVM SHA256-bit shape at k=19 (BASELINE config 3) and the aggregation shape at k=20 (config 5 reduced one notch; k=22 needs
> 1 h of oracle time).  Run once per round on a GPU box; the result is committed under profiles/.

    python tools/parity_big.py [sha_k19 agg_k20 ...]      ->  gpurun_out/r02_parity_big.json
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
ctx = pkg.Context(0)
names = sys.argv[1:] or ["sha_k19", "agg_k20"]
out = {}
for name in names:
    wl = bench.WORKLOADS[name]
    t0 = time.perf_counter()
    w = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100, shape=wl.get("shape", "base"))
    seed = pkg.seed_from_u64(4242)
    gpu = pkg.create_proof(w.pk, w.advice_dev, w.instances, seed)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    prove, cores = bench.oracle_prover(pkg, w.circ)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu = prove(seed)
    t_cpu = time.perf_counter() - t0
    rec = {"desc": wl["desc"], "k": wl["k"], "proof_bytes": len(gpu), "gpu_sha256": hashlib.sha256(gpu).hexdigest(),
           "oracle_sha256": hashlib.sha256(cpu).hexdigest(), "bytes_equal": gpu == cpu, "oracle_create_proof_s": round(t_cpu, 2),
           "oracle_setup_s": round(t_setup, 2), "oracle_cores": cores, "gpu_build_and_prove_s": round(t_gpu, 2)}
    if w.compact is not None:
        rec["compact_equal"] = pkg.create_proof_compact(w.pk, w.compact, w.instances, seed) == cpu
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    del w
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity_big.json"), "w"), indent=1)
assert all(r["bytes_equal"] for r in out.values())
