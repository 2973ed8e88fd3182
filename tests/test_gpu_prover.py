"""GPU parity: zkc_prove through the C ABI vs the CPU oracle's create_proof — proof bytes identical for
the same pk, witness, ChaCha20 seed and transcript; and the proofs verify under the independent verifier."""
import numpy as np
import pytest

from oracle import orc, plonk, verifier
from tests import pyref
from tests.circuits import SRS_SECRET, cols_to_mont, oracle_setup, oracle_srs
from tests.util import gpu_ctx, pkg

pytestmark = pytest.mark.gpu


def gpu_setup(circ, opk, zeta_choice=0):
    p = pkg()
    ctx = gpu_ctx()
    cs = circ.cs
    g, gl = oracle_srs(cs.k)
    params = p.ParamsKZG(cs.k, g=g, g_lagrange=gl, ctx=ctx)
    fixed = np.concatenate(opk.fixed_values) if opk.fixed_values else np.zeros((0, 4), dtype=np.uint64)
    sigma = np.concatenate(opk.sigma_values) if opk.sigma_values else np.zeros((0, 4), dtype=np.uint64)
    gpk = p.ProvingKey(params, cs, fixed, sigma, orc.fr_from_ints([opk.transcript_repr]), zeta_choice)
    return params, gpk


def vk_of(opk):
    return verifier.VerifyingKey(opk.cs, opk.fixed_commitments, opk.sigma_commitments, opk.transcript_repr)


@pytest.fixture(scope="module")
def circuit_k6():
    circ = pkg().synth.make_base_circuit(6, 2, seed=1)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    return circ, opk, advice, params, gpk


def test_pk_matches_oracle_keygen(circuit_k6):
    circ, opk, advice, params, gpk = circuit_k6
    assert (gpk.k, gpk.extended_k, gpk.degree, gpk.blinding_factors, gpk.num_sets, gpk.num_lookups) == (6, 8, 4, 6, 3, 1)
    f, s = gpk.commitments()
    assert orc.g1_to_ints(f) == opk.fixed_commitments
    assert orc.g1_to_ints(s) == opk.sigma_commitments


@pytest.mark.parametrize("transcript,multiopen", [("blake2b", "shplonk"), ("keccak", "shplonk"), ("blake2b", "gwc"), ("keccak", "gwc"),
                                                  ("evm", "shplonk"), ("evm", "gwc"), ("poseidon", "shplonk"), ("poseidon", "gwc")])
def test_proof_bytes_match_oracle(circuit_k6, transcript, multiopen):
    circ, opk, advice, params, gpk = circuit_k6
    seed = pyref.seed_from_u64(42)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20), transcript, multiopen)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    got = pkg().create_proof(gpk, np.concatenate(advice), inst, seed, transcript, multiopen)
    assert got == want
    assert verifier.verify_proof(vk_of(opk), pyref.G1_GEN, circ.instances, got, verifier.trapdoor_check(SRS_SECRET), transcript, multiopen)


@pytest.mark.parametrize("opts", [dict(advice_blinding="pse"), dict(blind_draws=True), dict(point_format=1),
                                  dict(advice_blinding="pse", blind_draws=True),
                                  dict(lookup_fill="axiom"),                                          # SURVEY OPEN-9
                                  dict(random_poly="chunked", random_poly_threads=4),                 # SURVEY OPEN-3
                                  dict(random_poly="chunked", random_poly_threads=3, blind_draws=True),
                                  dict(random_poly="chunked", random_poly_threads=1, blind_draws=True, advice_blinding="pse"),
                                  dict(random_poly="chunked", random_poly_threads=64, lookup_fill="axiom", blind_draws=True)])
def test_open_switches_match_oracle(circuit_k6, opts):
    circ, opk, advice, params, gpk = circuit_k6
    seed = pyref.seed_from_u64(7)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20), opts=plonk.ProverOptions(**opts))
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    got = pkg().create_proof(gpk, np.concatenate(advice), inst, seed, **opts)
    assert got == want


def test_rng_and_seed_helpers():
    assert pkg().seed_from_u64(0) == pyref.seed_from_u64(0)


@pytest.mark.parametrize("k,a,seed", [(6, 1, 3), (8, 3, 4), (10, 2, 5), (12, 3, 6)])
def test_proof_bytes_match_oracle_sizes(k, a, seed):
    circ = pkg().synth.make_base_circuit(k, a, seed=seed)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    s = pyref.seed_from_u64(seed)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(s, 20))
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    got = pkg().create_proof(gpk, np.concatenate(advice), inst, s)
    assert got == want
    assert verifier.verify_proof(vk_of(opk), pyref.G1_GEN, circ.instances, got, verifier.trapdoor_check(SRS_SECRET))


def test_device_resident_advice_and_errors(circuit_k6):
    import torch
    circ, opk, advice, params, gpk = circuit_k6
    seed = pyref.seed_from_u64(9)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    host = pkg().create_proof(gpk, np.concatenate(advice), inst, seed)
    dev = torch.from_numpy(np.concatenate(advice).view(np.int64)).cuda()
    assert pkg().create_proof(gpk, dev, inst, seed) == host
    # lookup input outside the table -> plonk::Error::ConstraintSystemFailure
    bad = [a.copy() for a in advice]
    bad[-1][3] = orc.fr_from_ints([1 << 40])[0]
    with pytest.raises(pkg().ZkcError) as e:
        pkg().create_proof(gpk, np.concatenate(bad), inst, seed)
    assert e.value.code == 11
    # too many instances -> InvalidInstances
    with pytest.raises(pkg().ZkcError) as e:
        pkg().create_proof(gpk, np.concatenate(advice), [orc.fr_from_ints([1] * 64)], seed)
    assert e.value.code == 10


def test_sha_bit_shape_matches_oracle():
    """BASELINE config-3 shape (many bit columns, ~75 small gates up to degree 4, rotation-1 queries, no
    lookup) at a size the oracle finishes in seconds."""
    circ = pkg().synth.make_sha_bit_circuit(9, 48, 3, blocks=4, seed=2)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    assert (gpk.degree, gpk.blinding_factors, gpk.num_sets, gpk.num_lookups) == (4, 5, 2, 0)
    s = pyref.seed_from_u64(21)
    for multiopen in ("shplonk", "gwc"):
        want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(s, 20), multiopen=multiopen)
        inst = [orc.fr_from_ints(c) for c in circ.instances]
        got = pkg().create_proof(gpk, np.concatenate(advice), inst, s, multiopen=multiopen)
        assert got == want
        assert verifier.verify_proof(vk_of(opk), pyref.G1_GEN, circ.instances, got, verifier.trapdoor_check(SRS_SECRET), multiopen=multiopen)


def test_wide_column_shape_matches_oracle():
    """BASELINE config-2 shape (12 gate columns + lookup: 15 permutation columns, 8 sets) at k=9."""
    circ = pkg().synth.make_base_circuit(9, 12, seed=8)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    assert gpk.num_sets == 8
    s = pyref.seed_from_u64(22)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(s, 20))
    got = pkg().create_proof(gpk, np.concatenate(advice), [orc.fr_from_ints(c) for c in circ.instances], s)
    assert got == want


def test_workload_builder_device_keygen_matches_oracle():
    """workload.build derives Montgomery columns, the sigma table and the SRS on the device; its proof must
    equal the oracle's built from the same synthetic circuit with host arithmetic."""
    ctx = gpu_ctx()
    w = pkg().workload.build(ctx, 8, 2, seed=3)
    opk, advice = oracle_setup(w.circ)
    f, sg = w.pk.commitments()
    assert orc.g1_to_ints(f) == opk.fixed_commitments and orc.g1_to_ints(sg) == opk.sigma_commitments
    s = pyref.seed_from_u64(5)
    assert pkg().create_proof(w.pk, w.advice_dev, w.instances, s) == plonk.create_proof(opk, advice, w.circ.instances, pyref.ChaChaRng(s, 20))
    assert pkg().create_proof(w.pk, w.advice_host, w.instances, s) == pkg().create_proof(w.pk, w.advice_dev, w.instances, s)


def _verify_workload(w, proof, **kw):
    """independent verifier on a GPU-built workload: vk commitments come from the device pk"""
    f, s = w.pk.commitments()
    vk = verifier.VerifyingKey(w.circ.cs, orc.g1_to_ints(f), orc.g1_to_ints(s), w.circ.transcript_repr())
    return verifier.verify_proof(vk, pyref.G1_GEN, w.circ.instances, proof, verifier.trapdoor_check(SRS_SECRET), **kw)


def _product_verify(w, proof, **kw):
    """the product's own host verifier (zkc_verify, C++ pairing) on a GPU-built workload: no oracle code involved"""
    api = pkg().api
    f, s = w.pk.commitments()
    s_g2 = api.g2_mul(api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    return api.verify_proof(w.circ.cs, f, s, orc.fr_from_ints([w.circ.transcript_repr()]), w.params.get_g(0)[:1], api.g2_generator(), s_g2,
                            w.instances, proof, **kw)


@pytest.mark.parametrize("transcript,multiopen", [("blake2b", "shplonk"), ("keccak", "gwc"), ("evm", "shplonk"), ("poseidon", "shplonk")])
def test_product_verifier_accepts_gpu_proofs(transcript, multiopen):
    """prove on the GPU, verify with zkc_verify on the host: the drop-in's create_proof -> verify_proof loop"""
    w = pkg().workload.build(gpu_ctx(), 10, 3, seed=12)
    seed = pyref.seed_from_u64(8)
    proof = pkg().create_proof(w.pk, w.advice_dev, w.instances, seed, transcript, multiopen)
    assert _product_verify(w, proof, transcript=transcript, multiopen=multiopen)
    assert _verify_workload(w, proof, transcript_kind=transcript, multiopen=multiopen)
    bad = bytearray(proof)
    bad[-5] ^= 2
    assert not _product_verify(w, bytes(bad), transcript=transcript, multiopen=multiopen)


@pytest.mark.parametrize("k,cols", [(17, 3), (15, 12)])
def test_full_size_proof_bytes_match_oracle(k, cols):
    """BASELINE.json configs 1 and 2 at their real sizes (RSA k=17 with 3 + 1 advice columns; k=15 with 12 + 1): the proof bytes
    of the device prover equal the CPU oracle's for the same circuit, witness and seed — also with the axiom-fork variants of
    the lookup fill order and the random polynomial switched on (SURVEY OPEN-9 / OPEN-3) and through GWC + Keccak."""
    ctx = gpu_ctx()
    circ = pkg().synth.make_base_circuit(k, cols, seed=11)
    opk, advice = oracle_setup(circ)
    w = pkg().workload.build(ctx, k, cols, circ=circ)
    f, sg = w.pk.commitments()
    assert orc.g1_to_ints(f) == opk.fixed_commitments and orc.g1_to_ints(sg) == opk.sigma_commitments
    seed = pyref.seed_from_u64(1000 + k)
    want = plonk.create_proof(opk, advice, circ.instances, orc.ChaCha20Rng(seed))
    got = pkg().create_proof(w.pk, w.advice_dev, w.instances, seed)
    assert got == want
    assert pkg().create_proof(w.pk, w.advice_host, w.instances, seed) == want
    assert _product_verify(w, got)
    o = plonk.ProverOptions(lookup_fill="axiom", random_poly="chunked", random_poly_threads=16)
    want2 = plonk.create_proof(opk, advice, circ.instances, orc.ChaCha20Rng(seed), "keccak", "gwc", opts=o)
    assert want2 != want
    assert pkg().create_proof(w.pk, w.advice_dev, w.instances, seed, "keccak", "gwc", lookup_fill="axiom", random_poly="chunked",
                              random_poly_threads=16) == want2


@pytest.mark.parametrize("k,cols,shape", [(17, 3, "base"), (15, 12, "base"), (15, 112, "sha_bit")])
def test_full_size_proofs_verify(k, cols, shape):
    """BASELINE.json sizes (config 1: RSA k=17; config 2: k=15 with 12 gate columns; config 3 shape at k=15) through the
    size-independent property the reference's own tests use — the proof verifies — plus determinism and
    device-resident == host-buffer.  (Byte parity with the oracle at these sizes: test_full_size_proof_bytes_match_oracle.)"""
    ctx = gpu_ctx()
    w = pkg().workload.build(ctx, k, cols, seed=11, shape=shape)
    seed = pyref.seed_from_u64(k)
    proof = pkg().create_proof(w.pk, w.advice_dev, w.instances, seed)
    assert _verify_workload(w, proof)
    assert _product_verify(w, proof)
    assert pkg().create_proof(w.pk, w.advice_host, w.instances, seed) == proof
    ctx.set_overlap(False)
    try:
        assert pkg().create_proof(w.pk, w.advice_dev, w.instances, seed) == proof     # stream overlap does not change bytes
    finally:
        ctx.set_overlap(True)
    bad = bytearray(proof)
    bad[40] ^= 4
    try:
        ok = _verify_workload(w, bytes(bad))
    except ValueError:
        ok = False
    assert not ok
    if shape == "base" and k == 17:
        gwc = pkg().create_proof(w.pk, w.advice_dev, w.instances, seed, transcript="keccak", multiopen="gwc")
        assert _verify_workload(w, gwc, transcript_kind="keccak", multiopen="gwc")


def test_compact_witness_upload_matches_full():
    """zkc_prove_compact (bit-packed / u16 columns expanded on the device) emits the same bytes as zkc_prove."""
    ctx = gpu_ctx()
    w = pkg().workload.build(ctx, 10, 48, seed=4, shape="sha_bit")
    assert w.compact is not None and w.compact.nbytes * 50 < w.advice_host.nbytes
    kinds = {k for k, _ in w.compact.columns}
    assert 1 in kinds            # bit-packed columns present
    seed = pyref.seed_from_u64(31)
    full = pkg().create_proof(w.pk, w.advice_host, w.instances, seed)
    assert pkg().create_proof_compact(w.pk, w.compact, w.instances, seed) == full
    assert _verify_workload(w, full)
    # mixed: one column passed as full Montgomery limbs
    n = 1 << 10
    cols = list(w.compact.columns)
    mixed = pkg().CompactAdvice([("fr", w.advice_host[:n])] + [({0: "fr", 1: "bits", 2: "u8", 3: "u16", 4: "u64"}[k], a) for k, a in cols[1:]])
    assert pkg().create_proof_compact(w.pk, mixed, w.instances, seed) == full


@pytest.mark.parametrize("k,with_perm", [(7, True), (9, True), (8, False)])
def test_generic_paths_multi_lookup(k, with_perm):
    """Two lookups (one theta-compressed over two expressions: full-width sort keys), degree-5 constraint system
    (chunk_len 3), negative rotation, Scaled / Constant nodes, two instance columns, and the no-permutation case."""
    circ = pkg().synth.make_multi_lookup_circuit(k, seed=k, with_permutation=with_perm)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    assert (gpk.degree, gpk.num_lookups, gpk.num_sets) == (5, 2, 1 if with_perm else 0)
    s = pyref.seed_from_u64(100 + k)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    for transcript, multiopen in (("blake2b", "shplonk"), ("keccak", "gwc")):
        want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(s, 20), transcript, multiopen)
        got = pkg().create_proof(gpk, np.concatenate(advice), inst, s, transcript, multiopen)
        assert got == want
        assert verifier.verify_proof(vk_of(opk), pyref.G1_GEN, circ.instances, got, verifier.trapdoor_check(SRS_SECRET), transcript, multiopen)
    # a pair that is not a table row -> ConstraintSystemFailure
    bad = [a.copy() for a in advice]
    bad[1][2] = orc.fr_from_ints([123456789])[0]
    with pytest.raises(pkg().ZkcError) as e:
        pkg().create_proof(gpk, np.concatenate(bad), inst, s)
    assert e.value.code == 11


def test_std_rng_chacha12(circuit_k6):
    """rng_kind = 1: rand 0.8 `StdRng` (ChaCha12), what snark-verifier-sdk constructs (SURVEY OPEN-6)"""
    circ, opk, advice, params, gpk = circuit_k6
    seed = pyref.seed_from_u64(77)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    want = plonk.create_proof(opk, advice, circ.instances, orc.ChaCha20Rng(seed, rounds=12), "poseidon")
    got = pkg().create_proof(gpk, np.concatenate(advice), inst, seed, "poseidon", rng="std")
    assert got == want and got != pkg().create_proof(gpk, np.concatenate(advice), inst, seed, "poseidon")


@pytest.mark.parametrize("blinding", ["axiom", "pse"])
def test_staged_witness_upload(circuit_k6, blinding):
    """large host witnesses travel on a copy stream in column groups (zkc_prove staged upload); forced here on a small circuit"""
    circ, opk, advice, params, gpk = circuit_k6
    seed = pyref.seed_from_u64(55)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    kw = dict(advice_blinding=blinding) if blinding != "axiom" else {}
    want = pkg().create_proof(gpk, np.concatenate(advice), inst, seed, **kw)
    ctx = gpu_ctx()
    try:
        ctx.set_tunable("stage_min_bytes", 1)
        assert pkg().create_proof(gpk, np.concatenate(advice), inst, seed, **kw) == want
        w = pkg().workload.build(ctx, 12, 7, seed=2)
        ctx.set_tunable("stage_min_bytes", -1)
        ref = pkg().create_proof(w.pk, w.advice_dev, w.instances, seed)
        ctx.set_tunable("stage_min_bytes", 1)
        assert pkg().create_proof(w.pk, w.advice_host, w.instances, seed) == ref
    finally:
        ctx.set_tunable("stage_min_bytes", -1)


def test_repeated_and_concurrent_proofs_are_stable(circuit_k6):
    """50 proofs leave device memory where it was (pool reuse, no leak); two host threads sharing one context
    (upstream calls from rayon workers) get serialised by the library and both obtain the right bytes."""
    import threading
    import torch
    circ, opk, advice, params, gpk = circuit_k6
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    adv = np.concatenate(advice)
    seeds = [pyref.seed_from_u64(1000 + i) for i in range(4)]
    want = [pkg().create_proof(gpk, adv, inst, s) for s in seeds]
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(50):
        assert pkg().create_proof(gpk, adv, inst, seeds[i % 4]) == want[i % 4]
    torch.cuda.synchronize()
    assert abs(torch.cuda.mem_get_info()[0] - free0) < (64 << 20)
    got = {}

    def work(tid):
        got[tid] = [pkg().create_proof(gpk, adv, inst, seeds[(tid + j) % 4]) for j in range(8)]
    ts = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for tid in range(2):
        assert got[tid] == [want[(tid + j) % 4] for j in range(8)]


def test_gpu_prover_reproduces_golden_fixtures():
    """tests/golden/proofs.json: the CUDA prover emits the committed bytes (no oracle prover run in this test; the oracle is
    only used to lay out the synthetic circuit's key material)"""
    import hashlib
    from tests import golden_util
    for name, entry in golden_util.load().items():
        circ = golden_util.circuit_of(entry)
        opk, advice = oracle_setup(circ)
        params, gpk = gpu_setup(circ, opk)
        f, s = gpk.commitments()
        assert orc.g1_to_ints(f) == golden_util.points_of(entry["fixed_commitments"])
        assert orc.g1_to_ints(s) == golden_util.points_of(entry["sigma_commitments"])
        inst = [orc.fr_from_ints(c) for c in circ.instances]
        for combo, rec in entry["proofs"].items():
            t, m = combo.split("/")
            proof = pkg().create_proof(gpk, np.concatenate(advice), inst, pyref.seed_from_u64(entry["rng_seed_u64"]), t, m)
            assert hashlib.sha256(proof).hexdigest() == rec["sha256"], (name, combo)
