"""halo2-zkcert_b200 — B200-native (sm_100a) halo2-axiom `create_proof` hot path behind a C ABI.

Python here is the test/bench harness and a thin mirror of the reference's operator names
(`best_fft`, `best_multiexp`, `EvaluationDomain`, `ParamsKZG`, `create_proof`) over ctypes; the
product is `libzkcert_cuda.so` (see include/zkcert_cuda.h).  There is no CPU fallback: every
compute entry point raises `ZkcError` when no CUDA device is present.

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()` (or
`importlib`), which registers it as the module `halo2_zkcert_b200`.
"""
from . import api, circuit, dist, synth, workload  # noqa: F401
from .api import (CompactAdvice, ProverSession, RandomPolySpec, ZkcError, Context, create_proof_compact, EvaluationDomain, ParamsKZG, ProvingKey, best_fft, best_multiexp, create_proof,  # noqa: F401
                  default_context, lib, lib_path, seed_from_u64, verify_proof)
