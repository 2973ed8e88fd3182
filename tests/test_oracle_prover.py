"""The oracle prover's proofs verify under the independent verifier restatement (trapdoor and real
pairing checks), for SHPLONK and GWC, Blake2b and Keccak transcripts, and every OPEN switch."""
import time

import pytest

from oracle import plonk, verifier, pairing
from tests import pyref
from tests.circuits import SRS_SECRET, fast_rng_for, oracle_setup, rng_for
from tests.util import pkg


def _vk(pk):
    return verifier.VerifyingKey(pk.cs, pk.fixed_commitments, pk.sigma_commitments, pk.transcript_repr)


@pytest.fixture(scope="module")
def small():
    circ = pkg().synth.make_base_circuit(6, 2, seed=1)
    pk, advice = oracle_setup(circ)
    return circ, pk, advice


@pytest.mark.parametrize("kind,multiopen", [("blake2b", "shplonk"), ("keccak", "shplonk"), ("blake2b", "gwc"), ("keccak", "gwc"), ("evm", "shplonk"), ("poseidon", "shplonk")])
def test_oracle_proof_verifies(small, kind, multiopen):
    circ, pk, advice = small
    proof = plonk.create_proof(pk, advice, circ.instances, rng_for(7), kind, multiopen)
    chk = verifier.trapdoor_check(SRS_SECRET)
    assert verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, proof, chk, kind, multiopen)
    # deterministic for a fixed seed; different seed -> different proof
    assert proof == plonk.create_proof(pk, advice, circ.instances, rng_for(7), kind, multiopen)
    assert proof != plonk.create_proof(pk, advice, circ.instances, rng_for(8), kind, multiopen)
    # tampering is rejected
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    try:
        ok = verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, bytes(bad), chk, kind, multiopen)
    except ValueError:
        ok = False
    assert not ok
    wrong_inst = [[(circ.instances[0][0] + 1) % 256] + circ.instances[0][1:]]
    assert not verifier.verify_proof(_vk(pk), pyref.G1_GEN, wrong_inst, proof, chk, kind, multiopen)


def test_oracle_proof_verifies_with_real_pairing(small):
    circ, pk, advice = small
    proof = plonk.create_proof(pk, advice, circ.instances, rng_for(3))
    s_g2 = pairing.g2_mul(pairing.G2_GEN, SRS_SECRET)
    assert verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, proof, verifier.pairing_check(s_g2))


@pytest.mark.parametrize("opts", [dict(advice_blinding="pse"), dict(blind_draws=True), dict(point_format=1),
                                  dict(advice_blinding="pse", blind_draws=True, zeta_choice=1)])
def test_open_switches(small, opts):
    circ, pk, advice = small
    if opts.get("zeta_choice"):
        pk, advice = oracle_setup(circ, zeta_choice=1)
    o = plonk.ProverOptions(**opts)
    proof = plonk.create_proof(pk, advice, circ.instances, rng_for(11), opts=o)
    assert verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, proof, verifier.trapdoor_check(SRS_SECRET),
                                 point_format=o.point_format)


@pytest.mark.parametrize("opts", [dict(lookup_fill="axiom"), dict(random_poly="chunked", random_poly_threads=4),
                                  dict(random_poly="chunked", random_poly_threads=3, blind_draws=True),      # 3 + 1 seeds: even, aligned
                                  dict(random_poly="chunked", random_poly_threads=1, blind_draws=True, advice_blinding="pse"),  # 1 seed: half a block off
                                  dict(lookup_fill="axiom", random_poly="chunked", random_poly_threads=8)])
def test_open9_open3_switches(small, opts):
    """SURVEY OPEN-9 (lookup fill order) and OPEN-3 (per-thread random polynomial): both variants give valid proofs that
    differ from the default ones; the pure-Python RNG and the C++ stream agree on the word-granular draws."""
    circ, pk, advice = small
    o = plonk.ProverOptions(**opts)
    proof = plonk.create_proof(pk, advice, circ.instances, rng_for(13), opts=o)
    assert verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, proof, verifier.trapdoor_check(SRS_SECRET))
    base = {k: v for k, v in opts.items() if k in ("blind_draws", "advice_blinding")}
    assert proof != plonk.create_proof(pk, advice, circ.instances, rng_for(13), opts=plonk.ProverOptions(**base))
    assert proof == plonk.create_proof(pk, advice, circ.instances, fast_rng_for(13), opts=o)


def test_lookup_fill_orders():
    """permute_expression_pair: same A' and the same multiset in S' for both fill orders; S'[row] = A'[row] on first occurrences;
    PSE writes the leftovers to the repeated rows from the last one backwards, the axiom variant forwards."""
    from oracle import orc

    class Cs:
        n = 16
        def blinding_factors(self):
            return 5
    U = 10
    inp = [3, 3, 3, 1, 1, 2, 2, 2, 2, 5]
    tab = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    pad = [0] * 6
    draw = iter(range(100, 200)).__next__
    a_p, s_p = plonk.permute_expression_pair(Cs(), orc.fr_from_ints(inp + pad), orc.fr_from_ints(tab + pad), draw, "pse")
    a_a, s_a = plonk.permute_expression_pair(Cs(), orc.fr_from_ints(inp + pad), orc.fr_from_ints(tab + pad), draw, "axiom")
    A = orc.fr_to_ints(a_p)[:U]
    assert A == sorted(inp) == orc.fr_to_ints(a_a)[:U]
    SP, SA = orc.fr_to_ints(s_p)[:U], orc.fr_to_ints(s_a)[:U]
    assert sorted(SP) == sorted(SA) == tab
    firsts = [i for i in range(U) if i == 0 or A[i] != A[i - 1]]
    reps = [i for i in range(U) if i not in firsts]
    assert all(SP[i] == A[i] == SA[i] for i in firsts)
    left = sorted(set(tab) - set(inp))
    assert [SA[i] for i in reps] == left and [SP[i] for i in reversed(reps)] == left


def test_zeta_choice_does_not_change_proof(small):
    """h(X) is unique whatever coset it is evaluated on."""
    circ, pk, advice = small
    pk1, advice1 = oracle_setup(circ, zeta_choice=1)
    assert plonk.create_proof(pk, advice, circ.instances, rng_for(5)) == plonk.create_proof(pk1, advice1, circ.instances, rng_for(5))


def test_unsatisfied_witness_fails(small):
    circ, pk, advice = small
    bad = [a.copy() for a in advice]
    # break one active gate cell: d of the first active gate in column 0
    row = next(r for r in range(0, pk.n, 4) if circ.fixed[0][r] == 1)
    bad[0][row + 3] = plonk.M(12345)[0]
    # like upstream, the prover does not notice; the proof it emits must be rejected
    proof = plonk.create_proof(pk, bad, circ.instances, rng_for(1))
    assert not verifier.verify_proof(_vk(pk), pyref.G1_GEN, circ.instances, proof, verifier.trapdoor_check(SRS_SECRET))


def test_lookup_value_outside_table_is_an_error(small):
    circ, pk, advice = small
    bad = [a.copy() for a in advice]
    bad[-1][3] = plonk.M(1 << 40)[0]
    with pytest.raises(ValueError, match="ConstraintSystemFailure"):
        plonk.create_proof(pk, bad, circ.instances, rng_for(1))


def test_fast_rng_gives_identical_proof(small):
    circ, pk, advice = small
    assert plonk.create_proof(pk, advice, circ.instances, rng_for(4)) == plonk.create_proof(pk, advice, circ.instances, fast_rng_for(4))


def test_poseidon_grain_known_answers_and_host_spec():
    """The Grain LFSR reproduces the published round constants of the (t = 3, R_F = 8, R_P = 57, BN254) instance
    (the same parameters circomlib / iden3 ship), and the product's C++ restatement generates the same spec."""
    from oracle import poseidon
    c, m = poseidon.spec()
    assert c[0] == [0x0ee9a592ba9a9518d05986d656f40c2114c4993c11bb29938d21d47304cd8e6e,
                    0x00f1445235f2148c5986587169fc1bcd887b08d4d00868df5696fff40956e864,
                    0x08dff3487e8ac99e1f29a058d0fa80b930c728730b7ab36ce879f3890ecf73f5]
    assert len(c) == 65
    hc, hm = pkg().api.poseidon_spec()
    assert hc == c and hm == m
    # sponge: duplex behaviour and padding rule
    s1, s2 = poseidon.PoseidonSponge(), poseidon.PoseidonSponge()
    s1.update([1, 2]); s2.update([1, 2, 0])
    assert s1.squeeze() != s2.squeeze()
    a = s1.squeeze()
    assert a != s1.squeeze()


def test_oracle_reproduces_golden_fixtures():
    """tests/golden/proofs.json (tools/make_golden.py): the oracle's keygen commitments and proof bytes have not drifted"""
    import hashlib
    from tests import golden_util
    from tests.circuits import oracle_setup
    for name, entry in golden_util.load().items():
        circ = golden_util.circuit_of(entry)
        opk, advice = oracle_setup(circ)
        assert opk.fixed_commitments == golden_util.points_of(entry["fixed_commitments"]), name
        assert opk.sigma_commitments == golden_util.points_of(entry["sigma_commitments"]), name
        for combo, rec in entry["proofs"].items():
            t, m = combo.split("/")
            proof = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(pyref.seed_from_u64(entry["rng_seed_u64"]), 20), t, m)
            assert len(proof) == rec["len"] and hashlib.sha256(proof).hexdigest() == rec["sha256"], (name, combo)
            if "hex" in rec:
                assert proof.hex() == rec["hex"]
