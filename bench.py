#!/usr/bin/env python
"""bench.py — `create_proof` seconds on the RSA-2048 k=17 shape (BASELINE.json config 1), one rank per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload rsa_k17|rsa_k15|sha_k19|agg_k22|...]
                    [--team] [--no-extras] [--no-cpu-baseline]

A step = one create_proof (SHPLONK, Blake2b transcript, seeded ChaCha20) of a synthetic, satisfying BaseConfig circuit: k=17,
3 gate advice columns + 1 lookup advice column, 1 lookup, 6 permutation columns (README.md:48 shape;
/root/reference/src/helpers.rs:97-172).  Independent proofs shard across GPUs with no collective (weak scaling; the cert
chain's proofs are independent — BASELINE config 4).

  value        seconds per proof with the witness already resident in HBM (whole job: step time / N)
  e2e          the same through the host-buffer C ABI: witness copied from pinned host memory inside the timed region,
               proof bytes returned to the host
  roofline     the dominant kernel (MSM bucket accumulation) against the MEASURED integer-pipe peak (IMAD.WIDE issue rate,
               profiles/r01_ffbench.json) — this path has no dense contraction and is not HBM-bound; `whole_msm_frac` is the same
               ratio over ALL msm.* phases; `roofline_other` is the NTT against the measured Montgomery-product rate
  cpu_baseline the restated CPU oracle (not halo2-axiom itself: no Rust toolchain in this image) at the REAL size of the
               workload on all host cores, and `parity.oracle_bytes_equal`: the last timed proof equals the oracle's bytes
  other_workloads (N = 1)  the other BASELINE configs in the same run, time-guarded: rsa_k15, sha_k19, agg_k22, and
               `throughput`: several proofs in flight on one GPU (one ctx + host thread each)
  team (N > 1) ONE proof over the N GPUs (strong scaling): agg_k22 and sha_k19, after the weak-scaling headline

`--impl reference` times the CPU oracle alone (rank 0 only) on the requested workload at its real size for the requested
steps; shapes whose oracle proof takes minutes (sha_k19, agg_k22) print value null rather than a scaled number.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "rsa_k17": dict(k=17, gate_cols=3, desc="RSA-2048 PKCS#1 verify shape: k=17, 3 gate advice + 1 lookup advice, 1 lookup, 6 permutation columns"),
    "rsa_k15": dict(k=15, gate_cols=12, desc="RSA-2048 wide-column shape: k=15, 12 gate advice + 1 lookup advice"),
    "rsa_k13": dict(k=13, gate_cols=3, desc="reduced smoke shape k=13"),
    "sha_k19": dict(k=19, gate_cols=112, shape="sha_bit", desc="zkEVM SHA256-bit shape: k=19, 112 bit advice + 3 word advice columns, 187 gates, no lookup"),
    "agg_k22": dict(k=22, gate_cols=17, shape="base_fast", desc="X509 aggregation shape: BaseConfig k=22, 17 gate advice + 1 lookup advice (lookup_bits 21), 20 permutation columns"),
    "agg_k20": dict(k=20, gate_cols=17, shape="base_fast", desc="aggregation shape reduced to k=20"),
    "sha_k15": dict(k=15, gate_cols=112, shape="sha_bit", desc="SHA256-bit shape reduced to k=15"),
}
# BASELINE config 4: the four independent proofs of a 3-certificate chain (/root/reference/src/tests/x509_aggregation.rs:34-57,
# src/bin/cli.rs:385-390), with the single-GPU time and the non-scaling share of each shape measured on this pool's B200s
# (profiles/r02_*): what the scheduler (dist.plan_chain) cuts the GPUs into teams with
CHAIN4 = ["rsa_k17", "sha_k19", "rsa_k17", "sha_k19"]
CHAIN_EST = {"rsa_k17": (0.0104, 0.75), "sha_k19": (0.0953, 0.15)}
ORACLE_MAX_K = {"base": 17, "base_fast": 17, "sha_bit": 15}   # real-size oracle proofs that finish within ~10 s on the box's cores
FQMUL_PER_MADD = 10      # XYZZ mixed add: 8M + 2S
IMADW_PER_FQMUL = 128    # 64 (a*b) + 64 (m*p) IMAD.WIDE per Montgomery product
EST_EXTRA_S = {"rsa_k15": 15, "throughput": 15, "sha_k19": 60, "agg_k22": 150}   # setup + a few steps, measured on the pool's boxes


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "imadw": 8.63e12, "imadw_src": "profiles/r01_ffbench.json"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"], peaks["hbm_src"] = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    for name in ("r02_msm_accum_traffic.json", "r01_msm_accum_traffic.json"):
        try:
            peaks["accum_traffic"] = json.load(open(os.path.join(ROOT, "profiles", name)))["avg_traffic_bytes"]
            peaks["accum_traffic_src"] = "profiles/" + name
            break
        except Exception:
            peaks["accum_traffic"] = None
    try:
        fb = json.load(open(os.path.join(ROOT, "profiles", "r01_ffbench.json")))
        peaks["imadw"] = float(fb["imad_wide_per_s"])
        peaks["fe_mul"] = float(fb["fr_mul_per_s"])
    except Exception:
        pass
    return peaks


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ---- CPU oracle (checker + reported baseline; never the product path) -----------------------------------------------
def make_circuit(pkg, wl, seed):
    gen = {"sha_bit": pkg.synth.make_sha_bit_circuit, "base_fast": pkg.synth.make_base_circuit_fast, "base": pkg.synth.make_base_circuit}[wl.get("shape", "base")]
    return gen(wl["k"], wl["gate_cols"], seed=seed)


def oracle_prover(pkg, circ):
    """(prove(seed32) -> proof bytes, cores) for the restated CPU oracle on this circuit; setup (SRS, keygen) is untimed"""
    from oracle import orc, plonk
    from tests import pyref
    cs = circ.cs
    g, gl = orc.srs_setup(cs.k, orc.fr_from_ints([pyref.ChaChaRng(bytes(32), 20).fr_random()]))
    if isinstance(circ.copies, __import__("numpy").ndarray):
        mapping = pkg.synth.build_permutation_mapping_fast(cs, circ.copies)
    else:
        mapping = pkg.synth.build_permutation_mapping(cs, [tuple(int(v) for v in c) for c in circ.copies])
    sigma = pkg.synth.sigma_values(cs, mapping)
    mont = lambda cols: [orc.fr_from_ints(c) for c in cols]
    if hasattr(circ, "advice_limbs"):
        fixed = [orc.field_op("fr", "from_canonical", c) for c in circ.fixed_limbs]
        advice = [orc.field_op("fr", "from_canonical", c) for c in circ.advice_limbs]
    else:
        fixed, advice = mont(circ.fixed), mont(circ.advice)
    pk = plonk.keygen(cs, fixed, mont(sigma), g, gl, circ.transcript_repr())
    return (lambda seed: plonk.create_proof(pk, advice, circ.instances, orc.ChaCha20Rng(seed))), orc.lib().orc_default_threads()


def time_oracle(prove, seeds, warmup):
    """seconds per proof (mean over seeds[warmup:]) and the last proof"""
    times, proof = [], None
    for i, s in enumerate(seeds):
        t0 = time.perf_counter()
        proof = prove(s)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), proof


def oracle_feasible(wl):
    return wl["k"] <= ORACLE_MAX_K[wl.get("shape", "base")]


def run_reference(args, wl):
    """the reference arm: the restated CPU oracle at the workload's REAL size, `steps` timed proofs after min(warmup, 1)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as graft
    pkg = graft.load_package()
    base = {"metric": "create_proof_s", "unit": "s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (BN254 Fr/Fq, exact)", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "k": wl["k"], "transcript": "blake2b", "multiopen": "shplonk"}}
    if not oracle_feasible(wl):
        why = "the restated CPU oracle needs minutes per proof at this size; no scaled number is reported"
        base.update({"value": None, "ms_per_step": None, "extrapolated": False, "unavailable": why,
                     "cpu_baseline": {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port", "sample": why},
                     "e2e": {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return 0
    circ = make_circuit(pkg, wl, seed=100)
    prove, cores = oracle_prover(pkg, circ)
    W = min(args.warmup, 1)
    seeds = [pkg.seed_from_u64(i) for i in range(W + args.steps)]
    t_wall = time.perf_counter()
    sec, _ = time_oracle(prove, seeds, W)
    t_wall = time.perf_counter() - t_wall
    sample = "restated CPU oracle (oracle/plonk.py + libzkc_oracle.so, not halo2-axiom) create_proof at the real size k=%d, same circuit " \
             "shape and witness generator as the GPU arm; mean of %d proofs after %d warm-up" % (wl["k"], args.steps, W)
    base.update({"value": sec, "ms_per_step": sec * 1e3, "wall_s_timed_region": t_wall,
                 "cpu_baseline": {"value": sec, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
                 "e2e": {"value": sec, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))
    return 0


# ---- GPU arm ---------------------------------------------------------------------------------------------------------
class Env:
    pass


def timed_proofs(env, w, advice, seeds, ctx=None):
    """one create_proof per seed, each bracketed by CUDA events on the launching stream, L2 flushed before each; returns (ms, proofs)"""
    import torch
    pkg = env.pkg
    total_ms, proofs = 0.0, []
    for s in seeds:
        env.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if isinstance(advice, pkg.CompactAdvice):
            proof = pkg.create_proof_compact(w.pk, advice, w.instances, s)
        else:
            proof = pkg.create_proof(w.pk, advice, w.instances, s)
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        proofs.append(proof)
    return total_ms, proofs


def rooflines(env, prof, prof_ms):
    peaks = env.peaks
    accum = prof.get("msm.accum", {"ms": 0.0, "n": 0})
    madds = prof.get("count:msm.madds", {"n": 0})["n"]
    imadw = madds * FQMUL_PER_MADD * IMADW_PER_FQMUL
    ach = imadw / (accum["ms"] * 1e-3) if accum["ms"] else 0.0
    msm_ms = sum(v["ms"] for kx, v in prof.items() if kx.startswith("msm."))
    whole = imadw / (msm_ms * 1e-3) if msm_ms else 0.0
    roofline = {"bound": "int", "kernel": "k_msm_accum (XYZZ bucket accumulation)", "achieved": ach / 1e12, "peak": peaks["imadw"] / 1e12,
                "unit": "T IMAD.WIDE/s", "frac": ach / peaks["imadw"], "traffic": peaks.get("accum_traffic"),
                "traffic_note": "avg dram__bytes_read+write per k_msm_accum launch, ncu capture of the rsa_k17 workload (%s)" % peaks.get("accum_traffic_src"),
                "timing_note": "per-kernel CUDA-event times from a separate pass of K steps with stream overlap disabled",
                "launches": accum["n"], "avg_launch_ms": accum["ms"] / max(accum["n"], 1),
                "algorithmic": "mixed adds per launch x 10 Fq-mul x 128 IMAD.WIDE (SURVEY 8d); peak = measured IMAD.WIDE issue rate (%s)" % peaks["imadw_src"],
                "share_of_step": accum["ms"] / prof_ms if prof_ms else None,
                "whole_msm_frac": whole / peaks["imadw"], "whole_msm_ms_per_launch_set": msm_ms / max(accum["n"], 1),
                "whole_msm_note": "the same mixed-add count over ALL msm.* phases (digits, sort, accumulate, bucket sums, reduction)"}
    ntt_ms = sum(prof.get(kx, {"ms": 0.0})["ms"] for kx in ("ntt.strided", "ntt.last"))
    ntt_n = sum(prof.get(kx, {"n": 0})["n"] for kx in ("ntt.strided", "ntt.last"))
    ntt_muls = prof.get("count:ntt.muls", {"n": 0})["n"]
    ntt_bytes = prof.get("count:ntt.bytes", {"n": 0})["n"]
    fe_peak = peaks.get("fe_mul", 65.14e9)
    roofline_ntt = {"bound": "int", "kernel": "k_ntt_strided + k_ntt_last (radix-2^s passes of best_fft / coset conversions)",
                    "achieved": (ntt_muls / (ntt_ms * 1e-3) / 1e9) if ntt_ms else 0.0, "peak": fe_peak / 1e9, "unit": "G Fr-mul/s",
                    "frac": (ntt_muls / (ntt_ms * 1e-3) / fe_peak) if ntt_ms else 0.0, "traffic": None,
                    "launches": ntt_n, "avg_launch_ms": ntt_ms / max(ntt_n, 1),
                    "algorithmic": "field products per launch (N/2 per butterfly stage + inter-pass twiddles + fused scalings) / CUDA-event time; "
                                   "peak = measured Montgomery-product rate (profiles/r01_ffbench.json: 128 IMAD.WIDE each)",
                    "hbm_gbs": ntt_bytes / (ntt_ms * 1e-3 or 1) / 1e9, "hbm_frac": ntt_bytes / (ntt_ms * 1e-3 or 1) / (peaks["hbm_gbs"] * 1e9),
                    "share_of_step": ntt_ms / prof_ms if prof_ms else None}
    return (roofline, roofline_ntt) if accum["ms"] >= ntt_ms else (roofline_ntt, roofline), msm_ms, ntt_ms


def measure(env, name, K, W, team=False, seed_base=0, profile=True, e2e=True, keep=False):
    """build the workload, time K resident proofs and K end-to-end proofs (after W warm-ups); returns a record dict"""
    import torch
    pkg, ctx, wl = env.pkg, env.ctx, WORKLOADS[name]
    t_setup = time.perf_counter()
    w = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100 + (0 if team else env.rank), shape=wl.get("shape", "base"))
    t_setup = time.perf_counter() - t_setup
    seeds = [pkg.seed_from_u64(1000 * (0 if team else env.rank) + seed_base + i) for i in range(W + K)]
    timed_proofs(env, w, w.advice_dev, seeds[:W])
    env.barrier()
    l0 = ctx.launches
    t_wall = time.perf_counter()
    dev_ms, proofs_dev = timed_proofs(env, w, w.advice_dev, seeds[W:])
    env.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = ctx.launches - l0
    rec = {"workload": name, "desc": wl["desc"], "k": wl["k"], "extended_k": w.pk.extended_k, "steps": K, "warmup": W,
           "ms_per_step": env.max_over_ranks(dev_ms) / K, "gpu_launches_per_step": launches // K, "proof_bytes": len(proofs_dev[0]),
           "setup_s": round(t_setup, 2), "wall_s_timed_region": t_wall}
    if profile:
        # per-kernel CUDA-event timers (zkc_profile_*) over the same K steps, on the launching stream; kept out of the
        # headline loop because the extra event records perturb the host-side pacing
        ctx.set_overlap(False)     # one stream: each kernel's event time is its own, not shared with side-stream work
        ctx.profile_enable(True)
        ctx.profile_report()
        prof_ms, proofs_prof = timed_proofs(env, w, w.advice_dev, seeds[W:])
        prof = ctx.profile_report()
        ctx.profile_enable(False)
        ctx.set_overlap(True)
        assert proofs_prof == proofs_dev
        (rec["roofline"], rec["roofline_other"]), msm_ms, ntt_ms = rooflines(env, prof, prof_ms)
        rec["phases_ms_per_step"] = {kx: round(v["ms"] / K, 4) for kx, v in sorted(prof.items()) if not kx.startswith("count:")}
        rec["msm_points_per_s"] = prof.get("count:msm.points", {"n": 0})["n"] / (msm_ms * 1e-3 or 1)
        rec["ntt_ms_per_step"] = ntt_ms / K
    if e2e:
        # host (pinned) witness in, proof bytes out.  Bit / byte valued witnesses (SHA256-bit shape) cross PCIe in compact
        # form (zkc_prove_compact: the Rust host packs Assigned<Fr> cells into bit / u8 / u16 / u64 columns); the full Fr
        # columns and the cost of that packing are reported beside it
        forms = [("fr", w.advice_host, w.h2d_bytes)]
        if w.compact is not None:
            forms.insert(0, ("compact", w.compact, w.compact.nbytes + sum(i.nbytes for i in w.instances)))
        for form, host_witness, h2d in forms:
            timed_proofs(env, w, host_witness, seeds[:1])
            env.barrier()
            ms, proofs = timed_proofs(env, w, host_witness, seeds[W:])
            env.barrier()
            assert proofs == proofs_dev, "device-resident and host-buffer paths must emit identical proofs"
            r = {"value": env.max_over_ranks(ms) / K / 1e3, "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": len(proofs_dev[0]),
                 "witness_form": "compact (bit/u8/u16/u64 columns)" if form == "compact" else "Fr columns, pinned"}
            if form == forms[0][0]:
                rec["e2e"] = r
            else:
                rec["e2e_fr_columns"] = r
        if w.compact is not None:
            t0 = time.perf_counter()
            pkg.api.CompactAdvice([pkg.api.CompactAdvice.pack_canonical(c) for c in w.circ.advice_limbs])
            rec["e2e"]["host_packing_s"] = time.perf_counter() - t0
            rec["e2e"]["host_packing_note"] = "numpy packing of canonical limb columns into bit / u8 / u16 / u64 form, outside the timed region"
    rec["_w"], rec["_proofs"], rec["_seeds"] = w, proofs_dev, seeds[W:]
    if not keep:
        strip(rec)
        del w
        torch.cuda.empty_cache()
    return rec


def strip(rec):
    for kx in ("_w", "_proofs", "_seeds"):
        rec.pop(kx, None)
    return rec


def throughput(env, w, K, nctx):
    """`nctx` create_proofs in flight on ONE GPU: one zkc_ctx (own streams, own scratch) and one host thread each, sharing the
    resident SRS / proving key.  Wall-clock over K proofs per thread; every proof checked against the single-stream bytes."""
    import torch
    pkg = env.pkg
    ctxs = [pkg.Context(env.local) for _ in range(nctx)]
    seeds = [[pkg.seed_from_u64(50000 + 100 * t + i) for i in range(K + 1)] for t in range(nctx)]
    want = [[pkg.create_proof(w.pk, w.advice_dev, w.instances, s) for s in seeds[t][1:]] for t in range(nctx)]
    got = [[] for _ in range(nctx)]
    for t in range(nctx):
        pkg.create_proof(w.pk, w.advice_dev, w.instances, seeds[t][0], ctx=ctxs[t])    # warm-up: scratch arenas, twiddles
    torch.cuda.synchronize()
    gate = threading.Barrier(nctx + 1)

    def worker(t):
        gate.wait()
        for s in seeds[t][1:]:
            got[t].append(pkg.create_proof(w.pk, w.advice_dev, w.instances, s, ctx=ctxs[t]))
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(nctx)]
    for th in ths:
        th.start()
    gate.wait()
    t0 = time.perf_counter()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert got == want, "concurrent proofs differ from the single-stream ones"
    for c in ctxs:
        c.close()
    return {"contexts": nctx, "proofs": nctx * K, "wall_s": dt, "proofs_per_s": nctx * K / dt, "ms_per_proof_amortised": dt / (nctx * K) * 1e3}


def run_chain(env, args):
    """BASELINE config 4: wall-clock of the whole 4-proof chain on N GPUs.  The GPUs are cut into teams by the scheduler; each
    team proves its share of the chain one proof after the other, every proof spread over the team (zkc_team_*).  Parity:
    every team proof is compared byte for byte with the same proof made by one GPU alone."""
    import torch
    import torch.distributed as dist
    pkg, ctx, world, rank = env.pkg, env.ctx, env.world, env.rank
    jobs = [(name,) + CHAIN_EST[name] for name in CHAIN4]
    teams, est = pkg.dist.plan_chain(jobs, world)
    groups = [dist.new_group(list(range(f, f + sz))) if (world > 1 and sz > 1) else None for f, sz, _ in teams]   # same order on every rank
    mine = pkg.dist.team_of(teams, rank)
    first, size, js = teams[mine]
    if size > 1:
        ctx.team_init(groups[mine])
    ws = {}
    for j in js:
        wl = WORKLOADS[CHAIN4[j]]
        ws[j] = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100 + j, shape=wl.get("shape", "base"))
    W, K = max(args.warmup, 1), args.steps
    seed_of = lambda j, i: pkg.seed_from_u64(7000 + 100 * j + i)

    def chain_once(i):
        return {j: pkg.create_proof(ws[j].pk, ws[j].advice_dev, ws[j].instances, seed_of(j, i)) for j in js}
    for i in range(W):
        chain_once(i)
    total_ms, wall, proofs = 0.0, 0.0, None
    for i in range(K):
        env.flush.zero_()
        env.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        proofs = chain_once(W + i)
        e1.record()
        torch.cuda.synchronize()
        ms = env.max_over_ranks(e0.elapsed_time(e1))
        env.barrier()
        wall += time.perf_counter() - t0
        total_ms += ms
    launches = ctx.launches
    # parity: the same proofs by this GPU alone (untimed), and their single-GPU times for the "one proof per GPU" comparison
    if size > 1:
        ctx.team_leave()
    single_ms = {}
    ok = True
    if rank == first:
        for j in js:
            env.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pkg.create_proof(ws[j].pk, ws[j].advice_dev, ws[j].instances, seed_of(j, 0))
            e0.record()
            alone = pkg.create_proof(ws[j].pk, ws[j].advice_dev, ws[j].instances, seed_of(j, W + K - 1))
            e1.record()
            torch.cuda.synchronize()
            single_ms[j] = e0.elapsed_time(e1)
            ok = ok and alone == proofs[j]
    gathered = pkg.dist.gather_objects({"ok": ok, "single_ms": single_ms, "proofs": {j: len(p) for j, p in (proofs or {}).items()} if rank == first else {}})
    if rank == 0:
        single = {}
        for part in gathered:
            single.update(part["single_ms"])
        all_ok = all(part["ok"] for part in gathered)
        naive = max(single.values()) if world >= len(CHAIN4) else None
        line = {"metric": "chain4_wall_s", "value": total_ms / K / 1e3, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "u256 (BN254 Fr/Fq Montgomery, exact integer)", "data": "synthetic",
                "config": {"workload": "chain4", "desc": "3-certificate chain: 2 x RSA k=17 + 2 x SHA256-bit k=19 proofs (BASELINE config 4), witnesses resident",
                           "jobs": CHAIN4, "teams": [{"first_rank": f, "gpus": sz, "jobs": [CHAIN4[j] for j in jj]} for f, sz, jj in teams],
                           "scheduler_estimate_s": est, "l2": "flushed between steps (256 MiB memset, untimed)",
                           "parity": "every team proof == the same proof by one GPU alone: %s" % all_ok},
                "single_gpu_ms_per_proof": {"%d:%s" % (j, CHAIN4[j]): round(v, 3) for j, v in sorted(single.items())},
                "one_proof_per_gpu_ms": naive, "sum_of_single_gpu_ms": sum(single.values()),
                "host_wall_s_per_chain": wall / K, "gpu_launches": int(launches)}
        assert all_ok, "a team proof differs from the single-GPU proof"
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rsa_k17", choices=sorted(WORKLOADS) + ["chain4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_workloads / throughput (N = 1) and the team record (N > 1)")
    ap.add_argument("--extras-budget-s", type=float, default=240.0, help="wall-clock budget for the extra records; what does not fit is reported as skipped")
    ap.add_argument("--team", action="store_true",
                    help="N > 1: the HEADLINE is ONE proof spread over the N GPUs (MSM by point range, transforms by column, h(X) by row block; "
                         "strong scaling) instead of N independent proofs (default, weak scaling)")
    args = ap.parse_args()
    if args.workload == "chain4" and args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"metric": "chain4_wall_s", "value": None, "unit": "s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                              "warmup": args.warmup, "higher_is_better": False, "config": {"workload": "chain4"},
                              "unavailable": "the restated CPU oracle needs minutes per SHA256 k=19 proof; no scaled number is reported"}))
        return 0
    wl = WORKLOADS.get(args.workload)
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import __graft_entry__ as graft
    env = Env()
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    env.pkg = pkg = graft.load_package()
    env.ctx = ctx = pkg.Context(local)
    ctx.use_torch_stream()
    env.peaks = load_peaks()
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    W = max(args.warmup, 3)
    K = args.steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    env.barrier, env.max_over_ranks = barrier, max_over_ranks

    if args.workload == "chain4":
        rc = run_chain(env, args)
        if world > 1:
            dist.destroy_process_group()
        return rc
    team = args.team and world > 1
    if team:
        ctx.team_init()      # every rank proves the SAME certificate together (zkc_team_init: NCCL over NVLink); joined before
                             # the SRS is built so that its window tables are sized for the per-rank point range
    # ---- headline: untimed setup (gen_srs + gen_pk + witness; each rank proves its own certificate), then the timed regions
    sampler = ClockSampler(local)
    sampler.start()
    rec = measure(env, args.workload, K, W, team=team, keep=True)
    clocks = sampler.stop()
    w, proofs_dev, seeds = rec.pop("_w"), rec.pop("_proofs"), rec.pop("_seeds")
    # the timed proofs are real proofs: the product's host verifier (zkc_verify, no oracle code) accepts the last one
    verified = None
    parity = None
    line = None
    if rank == 0:
        api = pkg.api
        f_comm, s_comm = w.pk.commitments()
        s_g2 = api.g2_mul(api.g2_generator(), api.fr_random_stream(pkg.workload.GEN_SRS_SEED, 1))
        verified = bool(api.verify_proof(w.circ.cs, f_comm, s_comm, w.transcript_repr, w.params.get_g(0)[:1],
                                         api.g2_generator(), s_g2, w.instances, proofs_dev[-1]))
        assert verified, "zkc_verify rejected a timed proof"
        per = 1 if team else world     # proofs finished per step
        step_ms = rec["ms_per_step"]
        line = {"metric": "create_proof_s", "value": step_ms / 1e3 / per, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong" if team else "weak", "vs_baseline": None,
                "dtype": "u256 (BN254 Fr/Fq Montgomery, exact integer)", "data": "synthetic",
                "config": {"workload": args.workload, "desc": wl["desc"], "k": wl["k"], "extended_k": rec["extended_k"],
                           "proofs_per_step": per, "transcript": "blake2b", "multiopen": "shplonk",
                           "parallelism": ("team%d: one proof over %d GPUs (MSM by point range, coset transforms by residue class, h(X) by row block)" % (world, world))
                           if team else ("independent proofs, one per GPU" if world > 1 else "single GPU"),
                           "l2": "flushed between steps (256 MiB memset, untimed)", "proof_bytes": rec["proof_bytes"],
                           "proof_verified": verified},
                "e2e": dict(rec["e2e"], value=rec["e2e"]["value"] / per),
                "gpu_launches": int(rec["gpu_launches_per_step"] * K), "clocks": clocks,
                "roofline": rec["roofline"], "roofline_other": rec["roofline_other"],
                "roofline_hbm": {"bound": "hbm", "kernel": "k_ntt_strided + k_ntt_last", "note": "NTT passes are integer-pipe bound on 254-bit fields (DESIGN.md 4)",
                                 "achieved": (rec["roofline_other"] if "hbm_gbs" in rec["roofline_other"] else rec["roofline"]).get("hbm_gbs"),
                                 "peak": env.peaks["hbm_gbs"], "unit": "GB/s", "peak_src": env.peaks["hbm_src"]},
                "phases_note": "CUDA-event ms per step with stream overlap disabled (sum exceeds ms_per_step when overlap hides work)",
                "phases_ms_per_step": rec["phases_ms_per_step"], "msm_points_per_s": rec["msm_points_per_s"],
                "wall_s_timed_region": rec["wall_s_timed_region"]}
        if "e2e_fr_columns" in rec:
            line["e2e_fr_columns"] = dict(rec["e2e_fr_columns"], value=rec["e2e_fr_columns"]["value"] / per)
        # ---- CPU oracle at the real size: reported baseline + byte parity of the last timed proof (outside every timed region)
        if not args.no_cpu_baseline and world == 1:
            if oracle_feasible(wl):
                prove, cores = oracle_prover(pkg, w.circ)
                sec, oproof = time_oracle(prove, [seeds[0], seeds[0], seeds[-1]], 1)
                parity = {"oracle_bytes_equal": oproof == proofs_dev[-1], "checked": "last timed proof vs oracle.plonk.create_proof, same circuit / witness / seed"}
                line["cpu_baseline"] = {"value": sec, "unit": "s", "cores": cores, "kind": "port",
                                        "sample": "restated CPU oracle create_proof at the real size k=%d on the same circuit and witness, mean of 2 proofs "
                                                  "after 1 warm-up; published halo2-axiom figures for the RSA k=17 shape: 3.144 s (M1) / 1.813 s "
                                                  "(c6a.48xlarge), README.md:48" % wl["k"]}
                line["parity"] = parity
                assert parity["oracle_bytes_equal"], "the last timed proof differs from the CPU oracle's"
            else:
                line["cpu_baseline"] = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "the restated CPU oracle needs minutes per proof at this size; not run inside bench.py "
                                                  "(tools/parity_big.py runs it once per round: profiles/r02_parity_big.json)"}
    # ---- extras, time-guarded ------------------------------------------------------------------------------------------
    # watchdog: an extra record must never cost the headline (a rank-local failure inside a collective would otherwise hang)
    done = threading.Event()

    def watchdog():
        if not done.wait(args.extras_budget_s + 240):
            if rank == 0:
                line["extras_error"] = "watchdog: the extra records did not finish; headline printed without them"
                print(json.dumps(line), flush=True)
            os._exit(0)
    if not args.no_extras:
        threading.Thread(target=watchdog, daemon=True).start()
    t_extras = time.perf_counter()
    left = lambda: args.extras_budget_s - (time.perf_counter() - t_extras)
    if not args.no_extras and world == 1 and args.workload == "rsa_k17":
        others = {}
        try:
            if left() > EST_EXTRA_S["throughput"]:
                others["throughput_rsa_k17"] = {"single_stream_ms_per_proof": rec["ms_per_step"],
                                                "concurrent": [throughput(env, w, max(4, K // 2), nctx) for nctx in (2, 4)]}
        except Exception as e:   # an extra must never cost the headline
            others["throughput_rsa_k17"] = {"error": repr(e)[:300]}
        del w
        torch.cuda.empty_cache()
        for name in ("rsa_k15", "sha_k19", "agg_k22"):
            if left() < EST_EXTRA_S[name]:
                others[name] = {"skipped": "extras budget (%.0f s) exhausted" % args.extras_budget_s}
                continue
            try:
                r = measure(env, name, 3 if name != "rsa_k15" else 10, 3 if name == "rsa_k15" else 1)
                others[name] = {"value": r["ms_per_step"] / 1e3, "unit": "s", "steps": r["steps"], "warmup": r["warmup"], "e2e": r["e2e"],
                                "roofline": r["roofline"], "roofline_other": r["roofline_other"], "gpu_launches_per_step": r["gpu_launches_per_step"],
                                "setup_s": r["setup_s"], "desc": r["desc"], "phases_ms_per_step": r["phases_ms_per_step"]}
                if "e2e_fr_columns" in r:
                    others[name]["e2e_fr_columns"] = r["e2e_fr_columns"]
            except Exception as e:
                others[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        if rank == 0:
            line["other_workloads"] = others
    elif not args.no_extras and world > 1 and not team:
        # strong scaling of ONE proof over the N GPUs (BASELINE config 5, and config 3), after the weak-scaling headline
        del w
        torch.cuda.empty_cache()
        trec = {}
        try:
            ctx.team_init()
            for name in ("agg_k22", "sha_k19"):
                ok = torch.tensor([1.0 if left() > EST_EXTRA_S[name] else 0.0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # every rank takes the same decision
                if ok.item() < 1:
                    trec[name] = {"skipped": "extras budget (%.0f s) exhausted" % args.extras_budget_s}
                    continue
                r = measure(env, name, 3, 1, team=True, profile=(name == "agg_k22"), keep=True)
                tw, tproofs = r.pop("_w"), r.pop("_proofs")
                r.pop("_seeds", None)
                trec[name] = {"value": r["ms_per_step"] / 1e3, "unit": "s", "steps": 3, "e2e": r["e2e"], "setup_s": r["setup_s"], "desc": r["desc"],
                              "scaling": "strong", "n_gpus": world}
                if rank == 0:   # the team's proof is a real proof: the product's host verifier accepts it
                    api = pkg.api
                    fc, sc = tw.pk.commitments()
                    s_g2 = api.g2_mul(api.g2_generator(), api.fr_random_stream(pkg.workload.GEN_SRS_SEED, 1))
                    trec[name]["proof_verified"] = bool(api.verify_proof(tw.circ.cs, fc, sc, tw.transcript_repr, tw.params.get_g(0)[:1],
                                                                         api.g2_generator(), s_g2, tw.instances, tproofs[-1]))
                del tw, tproofs
                torch.cuda.empty_cache()
                if "phases_ms_per_step" in r:
                    trec[name]["phases_ms_per_step"] = r["phases_ms_per_step"]
            ctx.team_leave()
        except Exception as e:
            trec["error"] = repr(e)[:300]
        if rank == 0:
            line["team"] = trec
    done.set()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
