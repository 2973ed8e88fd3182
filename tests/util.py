"""Shared helpers for the test-suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def pkg():
    return graft.load_package()


def random_fr_mont(n, seed):
    """n uniformly random Fr elements as Montgomery limbs (n, 4) — fast path: random 256-bit values
    reduced with numpy-free Python only for small n; for large n we sample 253-bit canonical values,
    which are valid Montgomery residues of *some* element (every value < r is)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)   # < 2^252 < r: canonical
    return np.ascontiguousarray(a)


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)


_ctx = None


def gpu_ctx():
    global _ctx
    if _ctx is None:
        _ctx = pkg().Context(0)
    return _ctx
