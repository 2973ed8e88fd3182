"""Builds oracle-side proving keys and witnesses from the synthetic circuits (tests only)."""
import numpy as np

from oracle import orc, plonk
from tests import pyref
from tests.util import pkg

SRS_SECRET = pyref.ChaChaRng(bytes(32), 20).fr_random()      # gen_srs: ChaCha20Rng zero seed
_srs_cache = {}


def oracle_srs(k):
    if k not in _srs_cache:
        _srs_cache[k] = orc.srs_setup(k, orc.fr_from_ints([SRS_SECRET]))
    return _srs_cache[k]


def cols_to_mont(cols):
    return [orc.fr_from_ints(c) for c in cols]


def oracle_setup(circ, zeta_choice=0):
    """(pk, advice columns in Montgomery form) for the oracle prover"""
    synth = pkg().synth
    cs = circ.cs
    g, gl = oracle_srs(cs.k)
    mapping = synth.build_permutation_mapping(cs, circ.copies)
    sigma = synth.sigma_values(cs, mapping)
    pk = plonk.keygen(cs, cols_to_mont(circ.fixed), cols_to_mont(sigma), g, gl, circ.transcript_repr(), zeta_choice)
    return pk, cols_to_mont(circ.advice)


def rng_for(seed):
    return pyref.ChaChaRng(pyref.seed_from_u64(seed), 20)


def fast_rng_for(seed):
    """the oracle's C++ ChaCha20 stream (same draws as rng_for, pinned in test_oracle_ops.py)"""
    return orc.ChaCha20Rng(pyref.seed_from_u64(seed))
