"""CPU: the product's host-side verifier (zkc_verify: plonk::verify_proof with SHPLONK / GWC over KZG, BN254 pairing in
C++) against proofs made by the oracle prover and against the oracle's own Python verifier and pairing.  No GPU needed:
this is host logic of the drop-in (the self-check gen_snark_shplonk runs after create_proof, /root/reference/src/helpers.rs:233)."""
import numpy as np
import pytest

from oracle import orc, pairing, plonk, verifier
from tests import pyref
from tests.circuits import SRS_SECRET, oracle_setup
from tests.pyref import R_MOD
from tests.util import pkg


def g2_to_ints(arr):
    v = orc.fq_to_ints(np.ascontiguousarray(arr).reshape(-1, 4))
    return ((v[0], v[1]), (v[2], v[3]))


def g2_from_ints(pt):
    return orc.fq_from_ints([pt[0][0], pt[0][1], pt[1][0], pt[1][1]]).reshape(1, 16)


def test_g2_generator_and_scalar_mul_match_oracle():
    api = pkg().api
    assert g2_to_ints(api.g2_generator()) == pairing.G2_GEN
    for s in [1, 2, 3, 12345, R_MOD - 1, SRS_SECRET]:
        got = g2_to_ints(api.g2_mul(api.g2_generator(), orc.fr_from_ints([s])))
        assert got == pairing.g2_mul(pairing.G2_GEN, s)
    # published vector: EIP-197's G2 generator is the point above; [r]G2 = identity
    assert g2_to_ints(api.g2_mul(api.g2_generator(), orc.fr_from_ints([0]))) == ((0, 0), (0, 0))


def test_pairing_bilinearity_and_oracle_agreement():
    api = pkg().api
    a, b = 0x1234567890abcdef1234567890abcdef, 0xfedcba0987654321fedcba0987654321
    G = pyref.G1_GEN
    aG = orc.g1_from_ints([pyref.ec_mul(G, a)])
    n_abG = orc.g1_from_ints([pyref.ec_neg(pyref.ec_mul(G, a * b % R_MOD))])
    bad = orc.g1_from_ints([pyref.ec_neg(pyref.ec_mul(G, (a * b + 1) % R_MOD))])
    g2 = api.g2_generator()
    bG2 = api.g2_mul(g2, orc.fr_from_ints([b]))
    assert api.pairing_check(np.concatenate([aG, n_abG]), np.concatenate([bG2, g2]))           # e(aG, bH) e(-abG, H) = 1
    assert not api.pairing_check(np.concatenate([aG, bad]), np.concatenate([bG2, g2]))
    assert not api.pairing_check(aG, bG2)                                                       # a single non-trivial pairing is not 1
    assert api.pairing_check(np.zeros((1, 8), dtype=np.uint64), g2)                             # identity contributes 1
    # the oracle's independent (py_ecc-style) pairing agrees on both outcomes
    assert pairing.pairing_product_is_one([(pyref.ec_mul(G, a), g2_to_ints(bG2)), (pyref.ec_neg(pyref.ec_mul(G, a * b % R_MOD)), pairing.G2_GEN)])
    with pytest.raises(pkg().ZkcError):
        api.pairing_check(orc.g1_from_ints([(1, 3)]), g2)                                        # not on the curve


@pytest.fixture(scope="module")
def setup_k6():
    circ = pkg().synth.make_base_circuit(6, 2, seed=1)
    opk, advice = oracle_setup(circ)
    api = pkg().api
    s_g2 = api.g2_mul(api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    return circ, opk, advice, s_g2


def _verify(circ, opk, s_g2, instances, proof, **kw):
    api = pkg().api
    return api.verify_proof(circ.cs, orc.g1_from_ints(opk.fixed_commitments), orc.g1_from_ints(opk.sigma_commitments),
                            orc.fr_from_ints([opk.transcript_repr]), orc.g1_from_ints([pyref.G1_GEN]), api.g2_generator(), s_g2,
                            [orc.fr_from_ints(c) for c in instances], proof, **kw)


@pytest.mark.parametrize("transcript,multiopen", [("blake2b", "shplonk"), ("keccak", "shplonk"), ("blake2b", "gwc"), ("keccak", "gwc"),
                                                  ("evm", "shplonk"), ("evm", "gwc"), ("poseidon", "shplonk"), ("poseidon", "gwc")])
def test_product_verifier_accepts_oracle_proofs_and_rejects_tampering(setup_k6, transcript, multiopen):
    circ, opk, advice, s_g2 = setup_k6
    proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(3), 20), transcript, multiopen)
    kw = dict(transcript=transcript, multiopen=multiopen)
    assert _verify(circ, opk, s_g2, circ.instances, proof, **kw)
    # the oracle's Python verifier (trapdoor form of the same check) agrees
    vk = verifier.VerifyingKey(circ.cs, opk.fixed_commitments, opk.sigma_commitments, opk.transcript_repr)
    assert verifier.verify_proof(vk, pyref.G1_GEN, circ.instances, proof, verifier.trapdoor_check(SRS_SECRET), transcript, multiopen)
    # any flipped evaluation bit, a wrong instance, truncation, trailing bytes, the wrong SRS: rejected
    bad = bytearray(proof)
    bad[len(proof) - 40] ^= 1
    assert not _verify(circ, opk, s_g2, circ.instances, bytes(bad), **kw)
    wrong = [list(c) for c in circ.instances]
    wrong[0][0] = (wrong[0][0] + 1) % R_MOD
    assert not _verify(circ, opk, s_g2, wrong, proof, **kw)
    assert not _verify(circ, opk, s_g2, circ.instances, proof[:-1], **kw)
    assert not _verify(circ, opk, s_g2, circ.instances, proof + b"\x00", **kw)
    other = pkg().api.g2_mul(pkg().api.g2_generator(), orc.fr_from_ints([SRS_SECRET + 1]))
    assert not _verify(circ, opk, other, circ.instances, proof, **kw)


def test_product_verifier_other_shapes_and_point_format():
    p = pkg()
    api = p.api
    s_g2 = api.g2_mul(api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    for circ in (p.synth.make_multi_lookup_circuit(7, seed=7), p.synth.make_sha_bit_circuit(9, 48, 3, blocks=4, seed=2),
                 p.synth.make_multi_lookup_circuit(8, seed=8, with_permutation=False)):
        opk, advice = oracle_setup(circ)
        for multiopen in ("shplonk", "gwc"):
            proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(5), 20), multiopen=multiopen)
            assert _verify(circ, opk, s_g2, circ.instances, proof, multiopen=multiopen)
            bad = bytearray(proof)
            bad[7] ^= 0x10          # a commitment: either off the curve or a different point
            assert not _verify(circ, opk, s_g2, circ.instances, bytes(bad), multiopen=multiopen)
    circ = p.synth.make_base_circuit(6, 1, seed=4)
    opk, advice = oracle_setup(circ)
    proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(6), 20), opts=plonk.ProverOptions(point_format=1))
    assert _verify(circ, opk, s_g2, circ.instances, proof, point_format=1)
    assert not _verify(circ, opk, s_g2, circ.instances, proof, point_format=0)


def test_product_verifier_accepts_golden_proofs():
    """the committed fixture bytes (tests/golden/proofs.json) verify under zkc_verify against the committed vk commitments —
    no prover runs in this test"""
    from tests import golden_util
    api = pkg().api
    s_g2 = api.g2_mul(api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    n = 0
    for name, entry in golden_util.load().items():
        circ = golden_util.circuit_of(entry)
        f = orc.g1_from_ints(golden_util.points_of(entry["fixed_commitments"]))
        s = orc.g1_from_ints(golden_util.points_of(entry["sigma_commitments"]))
        for combo, rec in entry["proofs"].items():
            if "hex" not in rec:
                continue
            t, m = combo.split("/")
            ok = api.verify_proof(circ.cs, f, s, orc.fr_from_ints([circ.transcript_repr()]), orc.g1_from_ints([pyref.G1_GEN]), api.g2_generator(),
                                  s_g2, [orc.fr_from_ints(c) for c in circ.instances], bytes.fromhex(rec["hex"]), transcript=t, multiopen=m)
            assert ok, (name, combo)
            n += 1
    assert n >= 10


def test_params_file_round_trip_and_validation():
    """kzg_bn254_{k}.srs codec (zkc_params_write / zkc_params_read, RawBytes layout): round trip, size law, point validation"""
    from tests.circuits import oracle_srs
    api = pkg().api
    k = 6
    g, gl = oracle_srs(k)
    g2 = api.g2_generator()
    s_g2 = api.g2_mul(g2, orc.fr_from_ints([SRS_SECRET]))
    data = api.params_to_bytes(k, g, gl, g2, s_g2)
    assert len(data) == 4 + 2 * 64 * (1 << k) + 256 and data[:4] == (k).to_bytes(4, "little")
    assert data[4:68] == g[0].tobytes()                      # raw Montgomery limbs of g[0] = the generator
    k2, g_r, gl_r, g2_r, s_g2_r = api.params_from_bytes(data)
    assert k2 == k and np.array_equal(g_r, g) and np.array_equal(gl_r, gl) and np.array_equal(g2_r, g2) and np.array_equal(s_g2_r, s_g2)
    bad = bytearray(data)
    bad[4 + 64 * 5 + 3] ^= 1                                 # g[5] leaves the curve
    with pytest.raises(pkg().ZkcError):
        api.params_from_bytes(bytes(bad))
    assert api.params_from_bytes(bytes(bad), checked=False)[0] == k     # RawBytesUnchecked trusts the file
    with pytest.raises(pkg().ZkcError):
        api.params_from_bytes(data[:-1])
    bad = bytearray(data)
    bad[-1] ^= 0x40                                          # s_g2.y.c1 out of range / off the curve
    with pytest.raises(pkg().ZkcError):
        api.params_from_bytes(bytes(bad))


def test_verifier_rejects_random_mutations_without_crashing(setup_k6):
    """fuzz: random byte strings and 150 single-byte mutations of a valid proof are all rejected (never accepted, never a crash);
    malformed constraint-system blobs are an error, not a verdict"""
    circ, opk, advice, s_g2 = setup_k6
    rng = np.random.default_rng(7)
    for multiopen in ("shplonk", "gwc"):
        proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(11), 20), multiopen=multiopen)
        assert _verify(circ, opk, s_g2, circ.instances, proof, multiopen=multiopen)
        for _ in range(75):
            bad = bytearray(proof)
            pos = int(rng.integers(0, len(bad)))
            bad[pos] ^= int(rng.integers(1, 256))
            assert not _verify(circ, opk, s_g2, circ.instances, bytes(bad), multiopen=multiopen), pos
        for n in [0, 1, 31, 32, 33, 64, len(proof) // 2, len(proof) + 32]:
            assert not _verify(circ, opk, s_g2, circ.instances, rng.integers(0, 256, n, dtype=np.uint8).tobytes(), multiopen=multiopen)
    # truncated / corrupted constraint system: ZkcError (BAD_ARG), not a crash
    import ctypes as C
    api = pkg().api
    blob = circ.cs.serialize()
    ok = C.c_int(0)
    o = api.ProveOpts()
    one = orc.fr_from_ints([1])
    for cut in [0, 3, 8, 40, len(blob) // 2, len(blob) - 1]:
        b = blob[:cut]
        st = api.lib().zkc_verify(b, C.c_size_t(len(b)), None, None, api._hp(one), api._hp(orc.g1_from_ints([pyref.G1_GEN])), api._hp(api.g2_generator()),
                                  api._hp(s_g2), None, None, C.c_size_t(0), proof, C.c_size_t(len(proof)), C.byref(o), C.byref(ok))
        assert st != 0 and ok.value == 0


def test_instance_column_count_is_checked(setup_k6):
    """plonk::Error::InvalidInstances when instances.len() != cs.num_instance_columns: too few would make the C side read past
    the caller's arrays, extra columns would be silently ignored."""
    circ, opk, advice, s_g2 = setup_k6
    proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(3), 20))
    assert _verify(circ, opk, s_g2, circ.instances, proof)
    for bad in ([], list(circ.instances) + [[1, 2, 3]]):
        with pytest.raises(pkg().ZkcError) as e:
            _verify(circ, opk, s_g2, bad, proof)
        assert e.value.code == 10


def test_identity_commitments_in_the_verifying_key():
    """all selectors off: the vk's selector commitments are the point at infinity; the host verifier multiplies and adds them
    like any other commitment (g1_mul on the identity)."""
    circ = pkg().synth.make_base_circuit(6, 2, seed=5, fill=0.0)
    assert not any(circ.fixed[0]) and not any(circ.fixed[1])
    opk, advice = oracle_setup(circ)
    assert opk.fixed_commitments[0] is None and opk.fixed_commitments[1] is None
    api = pkg().api
    s_g2 = api.g2_mul(api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    for multiopen in ("shplonk", "gwc"):
        proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(4), 20), multiopen=multiopen)
        assert verifier.verify_proof(verifier.VerifyingKey(circ.cs, opk.fixed_commitments, opk.sigma_commitments, opk.transcript_repr),
                                     pyref.G1_GEN, circ.instances, proof, verifier.trapdoor_check(SRS_SECRET), multiopen=multiopen)
        assert _verify(circ, opk, s_g2, circ.instances, proof, multiopen=multiopen)
        bad = bytearray(proof); bad[7] ^= 1
        assert not _verify(circ, opk, s_g2, circ.instances, bytes(bad), multiopen=multiopen)
