#!/usr/bin/env python
"""bench.py — `create_proof` seconds on the RSA-2048 k=17 shape (BASELINE.json), one rank per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload rsa_k17|rsa_k15|...]

A step = one create_proof (SHPLONK, Blake2b transcript, seeded ChaCha20) of a synthetic, satisfying
BaseConfig circuit: k=17, 3 gate advice columns + 1 lookup advice column, 1 lookup, 6 permutation
columns (README.md:48 shape; /root/reference/src/helpers.rs:97-172).  Independent proofs shard across
GPUs with no collective (weak scaling; the cert chain's proofs are independent — BASELINE config 4).

  value  = seconds per proof with the witness already resident in HBM (whole job: step time / N)
  e2e    = the same through the host-buffer C ABI: witness copied from pinned host memory inside
           the timed region, proof bytes returned to the host
  roofline = the dominant kernel (MSM bucket accumulation) against the MEASURED integer-pipe peak
           (IMAD.WIDE issue rate, profiles/r01_ffbench.json) — this path has no dense contraction
           and is not HBM-bound; `roofline_hbm` gives the NTT kernels against the measured copy bandwidth.
  cpu_baseline = the restated CPU oracle (not halo2-axiom itself: no Rust toolchain in this image)
           on a bounded sample (k=15, same column shape), scaled by n*log2(n) to k=17.

`--impl reference` times that CPU oracle alone (rank 0 only).
"""
import argparse
import json
import math
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
SAMPLE_K = 17            # the CPU oracle runs the real k=17 shape (about 6-12 s per proof); larger shapes use k=12 scaled
IMAD_WIDE_PEAK = None    # filled from profiles/r01_ffbench.json
FQMUL_PER_MADD = 10      # XYZZ mixed add: 8M + 2S
IMADW_PER_FQMUL = 128    # 64 (a*b) + 64 (m*p) IMAD.WIDE per Montgomery product


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "imadw": 8.63e12, "imadw_src": "profiles/r01_ffbench.json"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"], peaks["hbm_src"] = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    try:
        peaks["accum_traffic"] = json.load(open(os.path.join(ROOT, "profiles", "r01_msm_accum_traffic.json")))["avg_traffic_bytes"]
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


def cpu_oracle_seconds(sample_k, gate_cols, steps=1, warmup=0, seed=1, shape="base"):
    """times the restated CPU oracle prover on the bounded sample; returns (seconds per proof, cores)"""
    import __graft_entry__ as graft
    pkg = graft.load_package()
    from oracle import orc, plonk
    from tests import pyref
    gen = {"sha_bit": pkg.synth.make_sha_bit_circuit, "base_fast": pkg.synth.make_base_circuit_fast, "base": pkg.synth.make_base_circuit}[shape]
    circ = gen(sample_k, gate_cols, seed=seed)
    cs = circ.cs
    g, gl = orc.srs_setup(sample_k, orc.fr_from_ints([pyref.ChaChaRng(bytes(32), 20).fr_random()]))
    mapping = pkg.synth.build_permutation_mapping(cs, [tuple(int(v) for v in c) for c in circ.copies])
    sigma = pkg.synth.sigma_values(cs, mapping)
    mont = lambda cols: [orc.fr_from_ints(c) for c in cols]
    pk = plonk.keygen(cs, mont(circ.fixed), mont(sigma), g, gl, circ.transcript_repr())
    advice = mont(circ.advice)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        plonk.create_proof(pk, advice, circ.instances, orc.ChaCha20Rng(pyref.seed_from_u64(i)))
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), orc.lib().orc_default_threads()


def scale_to(k_from, k_to):
    return (2 ** k_to * k_to) / (2 ** k_from * k_from)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    k = wl["k"]
    sk = min(SAMPLE_K, k)
    if wl.get("shape") in ("sha_bit", "base_fast"):
        sk = min(12, k)
    elif args.steps + min(args.warmup, 1) > 12:
        sk = min(15, k)      # keep the whole run within a few minutes
    sec, cores = cpu_oracle_seconds(sk, wl["gate_cols"], steps=args.steps, warmup=min(args.warmup, 1), shape=wl.get("shape", "base"))
    value = sec * scale_to(sk, k)
    sample = "restated CPU oracle (oracle/plonk.py + libzkc_oracle.so, not halo2-axiom) create_proof at k=%d, same column shape; " \
             "%.3f s measured%s" % (sk, sec, "" if sk == k else ", scaled by n*log2(n) to k=%d" % k)
    line = {"metric": "create_proof_s", "value": value, "unit": "s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (BN254 Fr/Fq, exact)", "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"], "k": k, "transcript": "blake2b", "multiopen": "shplonk"},
            "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rsa_k17", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--team", action="store_true",
                    help="N > 1: ONE proof spread over the N GPUs (MSM by point range, transforms by column, h(X) by row block; "
                         "strong scaling) instead of N independent proofs (default, weak scaling)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import numpy as np
    import torch
    import __graft_entry__ as graft
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    pkg = graft.load_package()
    ctx = pkg.Context(local)
    ctx.use_torch_stream()
    peaks = load_peaks()
    W = max(args.warmup, 3)
    K = args.steps

    # untimed setup: gen_srs + gen_pk + witness (each rank proves its own certificate: different seed)
    team = args.team and world > 1
    if team:
        ctx.team_init()      # every rank proves the SAME certificate together (zkc_team_init: NCCL over NVLink); joined before
                             # the SRS is built so that its window tables are sized for the per-rank point range
    w = pkg.workload.build(ctx, wl["k"], wl["gate_cols"], seed=100 + (0 if team else rank), shape=wl.get("shape", "base"))
    seeds = [pkg.seed_from_u64(1000 * (0 if team else rank) + i) for i in range(W + K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(advice, steps, warm):
        total_ms, proofs = 0.0, []
        for i in range(warm + steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if isinstance(advice, pkg.CompactAdvice):
                proof = pkg.create_proof_compact(w.pk, advice, w.instances, seeds[i])
            else:
                proof = pkg.create_proof(w.pk, advice, w.instances, seeds[i])
            e1.record()
            torch.cuda.synchronize()
            if i >= warm:
                total_ms += e0.elapsed_time(e1)
                proofs.append(proof)
        return total_ms, proofs

    # ---- device-resident timing -------------------------------------------------------------------
    run(w.advice_dev, 0, W)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launches
    t_wall = time.perf_counter()
    dev_ms, proofs_dev = run(w.advice_dev, K, 0)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = ctx.launches - l0
    # per-kernel CUDA-event timers (zkc_profile_*) over the same K steps, on the launching stream;
    # kept out of the headline loop because the extra event records perturb the host-side pacing
    ctx.set_overlap(False)     # one stream: each kernel's event time is its own, not shared with side-stream work
    ctx.profile_enable(True)
    ctx.profile_report()
    prof_ms, proofs_prof = run(w.advice_dev, K, 0)
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    ctx.set_overlap(True)
    clocks = sampler.stop()
    assert proofs_prof == proofs_dev
    # ---- end-to-end timing: host (pinned) witness in, proof bytes out -------------------------------
    # bit / byte valued witnesses (SHA256-bit shape) cross PCIe in compact form (zkc_prove_compact); others as full Fr columns
    host_witness = w.compact if w.compact is not None else w.advice_host
    h2d_bytes = (w.compact.nbytes + sum(i.nbytes for i in w.instances)) if w.compact is not None else w.h2d_bytes
    run(host_witness, 0, 1)
    barrier()
    e2e_ms, proofs_e2e = run(host_witness, K, 0)
    barrier()
    assert proofs_dev == proofs_e2e, "device-resident and host-buffer paths must emit identical proofs"

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # the timed proofs are real proofs: the product's host verifier (zkc_verify, no oracle code) accepts the last one
    verified = None
    if rank == 0:
        api = pkg.api
        f_comm, s_comm = w.pk.commitments()
        s_g2 = api.g2_mul(api.g2_generator(), api.fr_random_stream(pkg.workload.GEN_SRS_SEED, 1))
        verified = bool(api.verify_proof(w.circ.cs, f_comm, s_comm, w.transcript_repr, w.params.get_g(0)[:1],
                                         api.g2_generator(), s_g2, w.instances, proofs_dev[-1]))
        assert verified, "zkc_verify rejected a timed proof"
    dev_ms, e2e_ms = max_over_ranks(dev_ms), max_over_ranks(e2e_ms)
    step_ms, e2e_step_ms = dev_ms / K, e2e_ms / K
    if rank == 0:
        accum = prof.get("msm.accum", {"ms": 0.0, "n": 0})
        madds = prof.get("count:msm.madds", {"n": 0})["n"]
        imadw = madds * FQMUL_PER_MADD * IMADW_PER_FQMUL
        ach = imadw / (accum["ms"] * 1e-3) if accum["ms"] else 0.0
        roofline = {"bound": "int", "kernel": "k_msm_accum (XYZZ bucket accumulation)", "achieved": ach / 1e12, "peak": peaks["imadw"] / 1e12,
                    "unit": "T IMAD.WIDE/s", "frac": ach / peaks["imadw"], "traffic": peaks.get("accum_traffic"),
                    "traffic_note": "avg dram__bytes_read+write per k_msm_accum launch, ncu capture of this workload (profiles/r01_msm_accum_traffic.json)",
                    "timing_note": "per-kernel CUDA-event times from a separate pass of K steps with stream overlap disabled",
                    "launches": accum["n"], "avg_launch_ms": accum["ms"] / max(accum["n"], 1),
                    "algorithmic": "mixed adds per launch x 10 Fq-mul x 128 IMAD.WIDE (SURVEY 8d); peak = measured IMAD.WIDE issue rate (%s)" % peaks["imadw_src"],
                    "share_of_step": accum["ms"] / prof_ms if prof_ms else None}
        ntt_ms = sum(prof.get(kx, {"ms": 0.0})["ms"] for kx in ("ntt.strided", "ntt.last"))
        ntt_n = sum(prof.get(kx, {"n": 0})["n"] for kx in ("ntt.strided", "ntt.last"))
        ntt_muls = prof.get("count:ntt.muls", {"n": 0})["n"]
        fe_peak = peaks.get("fe_mul", 65.14e9)
        roofline_ntt = {"bound": "int", "kernel": "k_ntt_strided + k_ntt_last (radix-2^s passes of best_fft / coset conversions)",
                        "achieved": (ntt_muls / (ntt_ms * 1e-3) / 1e9) if ntt_ms else 0.0, "peak": fe_peak / 1e9, "unit": "G Fr-mul/s",
                        "frac": (ntt_muls / (ntt_ms * 1e-3) / fe_peak) if ntt_ms else 0.0, "traffic": None,
                        "launches": ntt_n, "avg_launch_ms": ntt_ms / max(ntt_n, 1),
                        "algorithmic": "field products per launch (N/2 per butterfly stage + inter-pass twiddles + fused scalings) / CUDA-event time; "
                                       "peak = measured Montgomery-product rate (profiles/r01_ffbench.json: 128 IMAD.WIDE each); HBM side: 64 B per "
                                       "element per pass = %.1f %% of the measured copy bandwidth"
                                       % (100.0 * prof.get("count:ntt.bytes", {"n": 0})["n"] / (ntt_ms * 1e-3 or 1) / (peaks["hbm_gbs"] * 1e9)),
                        "hbm_gbs": prof.get("count:ntt.bytes", {"n": 0})["n"] / (ntt_ms * 1e-3 or 1) / 1e9,
                        "share_of_step": ntt_ms / prof_ms if prof_ms else None}
        n, en = 1 << wl["k"], 1 << w.pk.extended_k
        per = 1 if team else world     # proofs finished per step
        line = {"metric": "create_proof_s", "value": step_ms / 1e3 / per, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong" if team else "weak", "vs_baseline": None,
                "dtype": "u256 (BN254 Fr/Fq Montgomery, exact integer)", "data": "synthetic",
                "config": {"workload": args.workload, "desc": wl["desc"], "k": wl["k"], "extended_k": w.pk.extended_k,
                           "proofs_per_step": per, "transcript": "blake2b", "multiopen": "shplonk",
                           "parallelism": ("team%d: one proof over %d GPUs (MSM by point range, transforms by column, h(X) by row block)" % (world, world))
                           if team else ("independent proofs, one per GPU" if world > 1 else "single GPU"),
                           "l2": "flushed between steps (256 MiB memset, untimed)", "proof_bytes": len(proofs_dev[0]),
                           "proof_verified": verified},
                "e2e": {"value": e2e_step_ms / 1e3 / per, "unit": "s", "h2d_bytes_per_step": int(h2d_bytes), "witness_form": "compact (bit/u8/u16/u64 columns)" if w.compact is not None else "Fr columns, pinned",
                        "d2h_bytes_per_step": len(proofs_dev[0])},
                "gpu_launches": int(launches), "clocks": clocks,
                # `roofline` = the kernel with the larger share of the step; the other one is kept beside it
                "roofline": roofline if accum["ms"] >= ntt_ms else roofline_ntt,
                "roofline_other": roofline_ntt if accum["ms"] >= ntt_ms else roofline,
                "roofline_hbm": {"bound": "hbm", "kernel": "k_ntt_strided + k_ntt_last", "achieved_gbs_note":
                                 "NTT passes are integer-pipe bound on 254-bit fields; see DESIGN.md", "ntt_ms_per_step": ntt_ms / K,
                                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_src": peaks["hbm_src"]},
                "phases_note": "CUDA-event ms per step with stream overlap disabled (sum exceeds ms_per_step when overlap hides work)",
                "phases_ms_per_step": {kx: round(v["ms"] / K, 4) for kx, v in sorted(prof.items()) if not kx.startswith("count:")},
                "msm_points_per_s": prof.get("count:msm.points", {"n": 0})["n"] / (sum(prof.get(kx, {"ms": 0.0})["ms"] for kx in prof if kx.startswith("msm.")) * 1e-3 or 1),
                "wall_s_timed_region": t_wall}
        if not args.no_cpu_baseline and world == 1:
            sk = min(12 if wl.get("shape") in ("sha_bit", "base_fast") else SAMPLE_K, wl["k"])
            sec, cores = cpu_oracle_seconds(sk, wl["gate_cols"], shape=wl.get("shape", "base"))
            line["cpu_baseline"] = {"value": sec * scale_to(sk, wl["k"]), "unit": "s", "cores": cores, "kind": "port",
                                    "sample": "restated CPU oracle create_proof at k=%d (same column shape), %.3f s measured%s; published halo2-axiom "
                                              "figures for the RSA k=17 shape: 3.144 s (M1) / 1.813 s (c6a.48xlarge), README.md:48"
                                              % (sk, sec, "" if sk == wl["k"] else ", scaled by n*log2(n) to k=%d" % wl["k"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
