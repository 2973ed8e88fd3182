"""CPU-side checks of the product's host logic: the C ABI library loads and exports every symbol
include/zkcert_cuda.h declares, fails loudly without a GPU, the constraint-system wire format and
A.5 numbers, the synthetic circuits (every gate / lookup / copy constraint satisfied), and the
multi-GPU sharding helpers under gloo with world_size 2."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from tests.util import ROOT, pkg
from tests.pyref import R_MOD


def test_library_exports_every_declared_symbol():
    p = pkg()
    lib = p.lib()
    header = open(os.path.join(ROOT, "include", "zkcert_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(zkc_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.zkc_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = pkg()
    with pytest.raises(p.ZkcError) as e:
        p.Context(0)
    assert e.value.code == 2
    with pytest.raises(p.ZkcError):
        p.best_fft(np.zeros((4, 4), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64), 2)


def test_host_rng_helpers_match_known_answers():
    p = pkg()
    assert p.seed_from_u64(0).hex() == "ecf273f981b5cd4587f0467306ad6cadd0d0a3e33317e767f29bea72d78a7dfe"
    s = p.api.fr_random_stream(bytes(32), 2)
    rinv = pow(1 << 256, -1, R_MOD)
    to_int = lambda row: sum(int(row[i]) << (64 * i) for i in range(4)) * rinv % R_MOD
    assert to_int(s[0]) == 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba   # gen_srs secret
    assert np.array_equal(p.api.fr_random_stream(bytes(32), 1, skip=1), s[1:2])
    from tests import pyref
    std = pyref.ChaChaRng(p.seed_from_u64(9), 12)
    got = p.api.fr_random_stream(p.seed_from_u64(9), 3, rng="std")
    assert [to_int(r) for r in got] == [std.fr_random() for _ in range(3)]


def test_constraint_system_numbers_and_wire_format():
    p = pkg()
    cs = p.synth.base_constraint_system(17, 3)
    assert (cs.degree(), cs.blinding_factors(), cs.usable_rows(), cs.permutation_chunk_len(), cs.num_permutation_sets()) == \
        (4, 6, (1 << 17) - 7, 2, 3)
    assert (cs.num_advice, cs.num_fixed, cs.num_instance, len(cs.advice_queries), len(cs.fixed_queries)) == (4, 5, 1, 13, 5)
    cs15 = p.synth.base_constraint_system(15, 12)
    assert cs15.num_permutation_sets() == 8 and len(cs15.permutation) == 15
    blob = cs.serialize()
    assert blob[:4] == b"ZKCS" and struct.unpack_from("<6I", blob, 4) == (1, 17, 4, 5, 1, 0)
    # postfix program of the FlexGate polynomial q*(a + b*c - d)
    words, consts = cs.program(cs.gates[0])
    ops = words[0::2]
    from halo2_zkcert_b200.circuit import OP_ADVICE, OP_FIXED, OP_MUL, OP_ADD, OP_NEG, OP_END
    assert ops == [OP_FIXED, OP_ADVICE, OP_ADVICE, OP_ADVICE, OP_MUL, OP_ADD, OP_ADVICE, OP_NEG, OP_ADD, OP_MUL, OP_END] and consts == []


@pytest.mark.parametrize("k,a", [(6, 2), (9, 3)])
def test_synthetic_circuit_is_satisfied(k, a):
    p = pkg()
    circ = p.synth.make_base_circuit(k, a, seed=5)
    cs, n = circ.cs, 1 << k
    U = cs.usable_rows()
    for i in range(a):
        col, sel = circ.advice[i], circ.fixed[i]
        for r in range(U):
            if sel[r]:
                assert (col[r] + col[r + 1] * col[r + 2] - col[r + 3]) % R_MOD == 0
    table = set(circ.fixed[a + 1][:U])
    assert all(v in table for v in circ.advice[a][:U])
    cols = {**{i: circ.advice[i] for i in range(a + 1)}, a + 1: circ.fixed[a], a + 2: circ.instances[0]}
    assert len(circ.copies) > n // 16
    for lc, lr, rc, rr in circ.copies:
        assert cols[lc][lr] == cols[rc][rr]
    # permutation mapping: a permutation whose cycles only join equal cells
    mapping = p.synth.build_permutation_mapping(cs, circ.copies)
    assert sorted(mapping.tolist()) == list(range(len(cs.permutation) * n))
    val = lambda idx: (cols[idx // n][idx % n] if idx % n < len(cols[idx // n]) else 0)
    moved = [i for i in range(len(mapping)) if mapping[i] != i]
    assert moved and all(val(i) == val(int(mapping[i])) for i in moved)


def _dist_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = pkg().dist
    jobs = ["rsa_cert3", "sha_cert3", "rsa_cert2", "sha_cert2", "agg"]
    proofs = d.prove_chain(jobs, lambda j: ("proof(%s)@%d" % (j, rank)).encode())
    total = d.sum_partials(rank + 1, lambda a, b: a + b)
    q.put((rank, proofs, total, d.shard_range(10, world, rank)))
    dist.destroy_process_group()


def test_sharding_helpers_gloo_world2():
    import torch.multiprocessing as mp
    d = pkg().dist
    assert [d.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert d.assign_round_robin(5, 2, 1) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    want = [b"proof(rsa_cert3)@0", b"proof(sha_cert3)@1", b"proof(rsa_cert2)@0", b"proof(sha_cert2)@1", b"proof(agg)@0"]
    assert res[0][1] == want and res[1][1] == want          # every rank holds the chain's proofs in job order
    assert res[0][2] == res[1][2] == 3
    assert [r[3] for r in res] == [(0, 5), (5, 10)]


def test_g1_sum_host_epilogue():
    """zkc_g1_sum is host-only: the point-sharded MSM epilogue can be checked without a GPU."""
    from oracle import orc
    from tests import pyref
    p = pkg()
    ks = [5, 7, R_MOD - 13]
    pts = orc.fixed_base_batch(orc.fr_from_ints(ks))
    one = orc.fq_from_ints([1])
    jac = np.concatenate([np.concatenate([pts[i:i + 1], one], axis=1) for i in range(3)])
    ident = np.concatenate([orc.fq_from_ints([0, 1, 0]).reshape(1, 12)])
    out = p.api.g1_sum(np.concatenate([jac, ident]))
    assert orc.g1_to_ints(out[:, :8])[0] == pyref.ec_mul(pyref.G1_GEN, sum(ks) % R_MOD)
    assert orc.fq_to_ints(p.api.g1_sum(np.zeros((0, 12), dtype=np.uint64)).reshape(3, 4)) == [0, 1, 0]
    # P + (-P) = identity
    neg = orc.g1_from_ints([pyref.ec_neg(orc.g1_to_ints(pts[0:1])[0])])
    two = np.concatenate([np.concatenate([pts[0:1], one], axis=1), np.concatenate([neg, one], axis=1)])
    assert orc.fq_to_ints(p.api.g1_sum(two).reshape(3, 4)) == [0, 1, 0]


def test_team_partition_arithmetic():
    """zkc_team_shard_range / zkc_team_classes (dist.cu): blocks tile the range; a rank's residue classes are the ones its
    block of the class-major extended domain overlaps (rotations never leave a class, so that is all it has to transform)."""
    import ctypes as C
    L = pkg().lib()
    for total, world in [(10, 3), (7, 8), (1 << 12, 4), (0, 2), (5, 5)]:
        prev = 0
        for r in range(world):
            lo, hi = C.c_uint64(), C.c_uint64()
            assert L.zkc_team_shard_range(C.c_uint64(total), world, r, C.byref(lo), C.byref(hi)) == 0
            assert lo.value == prev and hi.value - lo.value in (total // world, total // world + 1)
            prev = hi.value
        assert prev == total
    # residue classes: every class is covered, a rank's classes are exactly those its class-major row block overlaps
    for k, ncls, world in [(10, 4, 2), (10, 3, 4), (10, 3, 8), (10, 4, 3), (8, 7, 5), (6, 1, 2), (12, 2, 1)]:
        n, seen = 1 << k, set()
        for r in range(world):
            lo, hi, c0, c1 = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint32()
            L.zkc_team_shard_range(C.c_uint64(ncls * n), world, r, C.byref(lo), C.byref(hi))
            assert L.zkc_team_classes(k, ncls, world, r, C.byref(c0), C.byref(c1)) == 0
            want = {i // n for i in range(lo.value, hi.value)}
            assert set(range(c0.value, c1.value)) == want
            seen |= want
        assert seen == set(range(ncls))
    assert L.zkc_team_classes(12, 0, 2, 0, C.byref(C.c_uint32()), C.byref(C.c_uint32())) != 0
    assert L.zkc_team_shard_range(C.c_uint64(4), 2, 2, None, None) != 0


def test_host_hashes_match_published_vectors():
    """The product's own Blake2b-512 / Keccak-256 (csrc/host/hostutil.cpp; what Blake2bWrite / Keccak256Write hash with)
    against RFC 7693 Appendix A, the Keccak team's vectors and hashlib (personalised, multi-block, empty)."""
    import ctypes as C
    import hashlib
    L = pkg().lib()

    def h(kind, data, person=None):
        out = (C.c_uint8 * (64 if kind == 0 else 32))()
        buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
        pers = None if person is None else (C.c_uint8 * 16).from_buffer_copy(person)
        assert L.zkc_host_hash(kind, pers, buf, C.c_size_t(len(data)), out) == 0
        return bytes(out)
    assert h(0, b"abc").hex() == ("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
                                  "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923")          # RFC 7693 App. A
    assert h(1, b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert h(1, b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    person = b"Halo2-Transcript"
    rng = np.random.default_rng(0)
    for n in [0, 1, 63, 64, 127, 128, 129, 135, 136, 137, 255, 256, 1000, 4097]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert h(0, data, person) == hashlib.blake2b(data, digest_size=64, person=person).digest()
        assert h(0, data) == hashlib.blake2b(data, digest_size=64).digest()
    try:
        from Crypto.Hash import keccak as _k   # optional cross-check when pycryptodome exists
    except Exception:
        _k = None
    if _k is not None:
        for n in [135, 136, 137, 500]:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            assert h(1, data) == _k.new(digest_bits=256, data=data).digest()


def test_host_field_inversion_matches_bigint_and_fermat():
    """the host driver's fe_inv (binary extended Euclid, csrc/ff.cuh) against Python's pow(x, -1, p) and the Fermat ladder,
    for both fields: edge values and 2000 random elements"""
    import ctypes as C
    import random
    from oracle import orc
    from tests.pyref import P_MOD
    L = pkg().lib()
    rnd = random.Random(5)
    for field, mod, conv, back in ((0, R_MOD, orc.fr_from_ints, orc.fr_to_ints), (1, P_MOD, orc.fq_from_ints, orc.fq_to_ints)):
        vals = [0, 1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, 1 << 128, (1 << 253) - 1] + [rnd.randrange(mod) for _ in range(2000)]
        a = conv(vals)
        out0, out1 = np.zeros_like(a), np.zeros_like(a)
        assert L.zkc_host_fe_inv(field, 0, a.ctypes.data_as(C.c_void_p), out0.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals))) == 0
        assert L.zkc_host_fe_inv(field, 1, a.ctypes.data_as(C.c_void_p), out1.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals))) == 0
        assert np.array_equal(out0, out1)
        assert back(out0) == [pow(v, -1, mod) if v else 0 for v in vals]


def test_rust_bindings_are_in_sync_with_the_header():
    """integration/zkcert_cuda_sys.rs (tools/gen_rust_bindings.py) declares every function of include/zkcert_cuda.h with the
    committed text equal to a fresh generation"""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_gen_rs", os.path.join(root, "tools", "gen_rust_bindings.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(root, "integration", "zkcert_cuda_sys.rs")).read()
    assert committed == gen.generate()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "zkcert_cuda.h")).read(), flags=re.S)
    names = set(re.findall(r"\b(zkc_[a-z0-9_]+)\s*\(", header))
    declared = set(re.findall(r"pub fn (zkc_[a-z0-9_]+)\(", committed))
    assert names == declared
    assert "instances: *const *const Fr" in committed and "-> *const c_char" in committed


def test_c_client_links_and_runs_host_entry_points(tmp_path):
    """include/zkcert_cuda.h is plain C99 and libzkcert_cuda.so links from C: tests/cabi_host.c calls the host entry points and
    checks that the device ones fail loudly here (no GPU, no CPU fallback)"""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pkg().lib()
    libdir = os.path.join(root, "halo2-zkcert_b200")
    exe = str(tmp_path / "cabi_host")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", os.path.join(root, "tests", "cabi_host.c"), "-o", exe,
                           "-L" + libdir, "-lzkcert_cuda", "-Wl,-rpath," + libdir])
    import torch
    args = [exe] + (["--gpu"] if torch.cuda.is_available() else [])
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0 and "cabi_host ok" in out.stdout, out.stderr


def test_chain_scheduler_plans():
    """dist.plan_chain (BASELINE config 4): every GPU in exactly one team, every proof dealt exactly once, and the plans for the
    3-certificate chain (2 x RSA k=17 + 2 x SHA k=19) put the SHA proofs on the larger teams instead of one proof per GPU"""
    d = pkg().dist
    jobs = [("rsa", 0.0107, 0.75), ("sha", 0.122, 0.13), ("rsa", 0.0107, 0.75), ("sha", 0.122, 0.13)]
    for world in range(1, 9):
        teams, span = d.plan_chain(jobs, world)
        ranks = [r for f, s, _ in teams for r in range(f, f + s)]
        assert ranks == list(range(world))
        assert sorted(j for _, _, js in teams for j in js) == [0, 1, 2, 3]
        loads = [sum(d.team_time(jobs[j][1], s, jobs[j][2]) for j in js) for _, s, js in teams]
        assert abs(max(loads) - span) < 1e-12
        assert all(d.team_of(teams, r) == i for i, (f, s, _) in enumerate(teams) for r in range(f, f + s))
        if world >= 2:
            # never worse than one proof per GPU / round robin
            naive = max(sum(jobs[j][1] for j in range(r, 4, min(world, 4))) for r in range(min(world, 4)))
            assert span <= naive + 1e-12
    teams8, span8 = d.plan_chain(jobs, 8)
    assert [s for _, s, _ in teams8] == [4, 4] and all(sorted(jobs[j][0] for j in js) == ["rsa", "sha"] for _, _, js in teams8)
    teams4, _ = d.plan_chain(jobs, 4)
    assert [s for _, s, _ in teams4] == [2, 2]
    # equal jobs, as many GPUs as jobs: one each
    teams, _ = d.plan_chain([("a", 1.0, 0.5)] * 4, 4)
    assert [s for _, s, _ in teams] == [1, 1, 1, 1]


def pyref_mod():
    from tests import pyref
    return pyref.R_MOD


def test_gate_program_factoring_is_exact():
    """csrc/host/cs.h optimize_program (the stream the device interpreter runs): a leaf shared by a run of >= 3 constraints is
    taken out of the run.  The Horner fold over random query values equals the fold of the stream as parsed, for left and
    right factors, runs broken by other selectors / unfactorable forms / short runs, and many distinct run lengths."""
    import ctypes as C
    import random
    p = pkg()
    from oracle import orc
    cir = p.circuit
    rnd = random.Random(7)
    A = lambda i: ("advice", i)
    F = lambda i: ("fixed", i)
    I = lambda i: ("instance", i)
    P = lambda a, b: ("product", a, b)
    S = lambda a, b: ("sum", a, b)
    N = lambda a: ("neg", a)
    K = lambda v: ("const", v)

    def body(t):
        return [S(P(A(0), A(1)), N(A(2))), P(A(3), S(K(1), N(A(3)))), ("scaled", S(A(1), A(2)), 5 + t), S(P(P(A(0), A(2)), A(3)), K(t)), A(2)][t % 5]

    def check(polys, want_groups):
        naq, nfq, niq = 4, 3, 1
        cs = cir.ConstraintSystem(6, 4, 3, 1, [(c, 0) for c in range(naq)], [(c, 0) for c in range(nfq)], [(0, 0)], [polys], [], [])
        blob = cs.serialize()
        vals = lambda m: orc.fr_from_ints([rnd.randrange(pyref_mod()) for _ in range(m)])
        aq, fq, iq, mult = vals(naq), vals(nfq), vals(niq), vals(1)
        got = []
        for factored in (0, 1):
            out = np.zeros((1, 4), dtype=np.uint64)
            groups = C.c_uint32(99)
            rc = p.lib().zkc_host_fold_gates(blob, C.c_size_t(len(blob)), aq.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p),
                                             iq.ctypes.data_as(C.c_void_p), mult.ctypes.data_as(C.c_void_p), factored, out.ctypes.data_as(C.c_void_p),
                                             C.byref(groups))
            assert rc == 0
            got.append((out.copy(), groups.value))
        assert np.array_equal(got[0][0], got[1][0])
        assert got[0][1] == 0 and got[1][1] == want_groups, got[1][1]
        # and the value is the plain big-integer fold
        R = pyref_mod()
        env = {"advice": orc.fr_to_ints(aq), "fixed": orc.fr_to_ints(fq), "instance": orc.fr_to_ints(iq)}

        def ev(e):
            t = e[0]
            if t == "const": return e[1] % R
            if t in env: return env[t][e[1]]
            if t == "neg": return (-ev(e[1])) % R
            if t == "sum": return (ev(e[1]) + ev(e[2])) % R
            if t == "product": return ev(e[1]) * ev(e[2]) % R
            return ev(e[1]) * e[2] % R
        acc, m = 0, orc.fr_to_ints(mult)[0]
        for e in polys:
            acc = (acc * m + ev(e)) % R
        assert orc.fr_to_ints(got[1][0])[0] == acc

    check([P(F(0), body(t)) for t in range(7)], 1)                                             # one run, selector on the left
    check([P(body(t), F(1)) for t in range(5)], 1)                                             # ... on the right
    check([P(F(0), body(0)), P(F(0), body(1))], 0)                                             # too short
    check([P(F(0), body(t)) for t in range(3)] + [P(F(1), body(t)) for t in range(4)] + [body(2)] + [P(F(0), body(t)) for t in range(3)], 3)
    check([P(F(0), body(t)) for t in range(3)] + [S(F(0), body(1))] + [P(A(1), body(t)) for t in range(6)] + [P(I(0), body(3))], 2)
    check([P(F(0), P(F(0), body(t))) for t in range(4)], 1)                                    # nested: only the outer factor goes
    runs = []
    for L in range(3, 13):                                                                     # ten distinct lengths: two runs stay unfactored
        runs += [P(F(L % 3), body(t)) for t in range(L)] + [K(L)]
    check(runs, 8)
    check([P(S(F(0), K(1)), body(t)) for t in range(4)], 0)                                    # the shared factor is not a leaf
