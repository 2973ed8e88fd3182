"""Regenerates tests/golden/proofs.json: proof bytes of the CPU oracle prover for fixed synthetic circuits, seeds and
transcripts, plus the verifying-key commitments.  The fixtures anchor BOTH provers over time: the oracle must keep
reproducing them (tests/test_oracle_prover.py), the CUDA prover must emit the same bytes (tests/test_gpu_prover.py) and the
product's host verifier must accept them (tests/test_host_verify.py).  They are self-generated (the Rust reference cannot run
here and pins no bytes — DESIGN.md §2: parity unpinned), so they detect drift, not disagreement with halo2-axiom.

    python tools/make_golden.py          # rewrites tests/golden/proofs.json
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import plonk  # noqa: E402
from tests import pyref  # noqa: E402
from tests.circuits import oracle_setup  # noqa: E402
from tests.util import pkg  # noqa: E402

CASES = [
    # name, synth generator, args, kwargs, rng seed (seed_from_u64), [(transcript, multiopen)], store full bytes?
    ("base_k6", "make_base_circuit", (6, 2), dict(seed=1), 42,
     [(t, m) for t in ("blake2b", "keccak", "evm", "poseidon") for m in ("shplonk", "gwc")], True),
    ("multi_lookup_k7", "make_multi_lookup_circuit", (7,), dict(seed=7), 107, [("blake2b", "shplonk"), ("keccak", "gwc")], True),
    ("sha_bit_k9", "make_sha_bit_circuit", (9, 48, 3), dict(blocks=4, seed=2), 21, [("blake2b", "shplonk")], False),
    ("base_k10", "make_base_circuit", (10, 2), dict(seed=5), 5, [("blake2b", "shplonk")], False),
]


def build():
    out = {"note": "oracle-generated (tools/make_golden.py); sha256 is over the proof bytes", "cases": {}}
    synth = pkg().synth
    for name, gen, args, kw, seed, combos, full in CASES:
        circ = getattr(synth, gen)(*args, **kw)
        opk, advice = oracle_setup(circ)
        entry = {"generator": gen, "args": list(args), "kwargs": kw, "rng_seed_u64": seed,
                 "fixed_commitments": [[hex(v) for v in p] if p else None for p in opk.fixed_commitments],
                 "sigma_commitments": [[hex(v) for v in p] if p else None for p in opk.sigma_commitments], "proofs": {}}
        for t, m in combos:
            proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(seed), 20), t, m)
            rec = {"len": len(proof), "sha256": hashlib.sha256(proof).hexdigest()}
            if full:
                rec["hex"] = proof.hex()
            entry["proofs"]["%s/%s" % (t, m)] = rec
        out["cases"][name] = entry
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "proofs.json")
    json.dump(build(), open(path, "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
