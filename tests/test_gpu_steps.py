"""GPU parity of the STEP API (zkc_prove_begin ... zkc_prove_end; SURVEY 8b Tier B): the caller keeps the Fiat-Shamir
transcript and the RNG — here the oracle's Python transcripts and the pure-Python ChaCha restatement, playing the part of
the Rust host's `TranscriptWrite` and `rng` inside a patched `plonk::create_proof` (reference call path:
/root/reference/src/helpers.rs:233,299 -> gen_snark_shplonk -> create_proof) — and the library does the device work of each
round.  The bytes that transcript ends up holding must equal the oracle's proof, for every transcript kind, both multiopen
schemes and every OPEN switch."""
import numpy as np
import pytest

from oracle import orc, plonk
from oracle.orc import R_MOD
from tests import pyref
from tests.circuits import oracle_setup
from tests.test_gpu_prover import gpu_setup
from tests.util import pkg

pytestmark = pytest.mark.gpu

M = lambda ints: orc.fr_from_ints(list(ints))


def host_create_proof(gpk, transcript_repr, advice, instances, rng, transcript_kind="blake2b", multiopen="shplonk", opts=None,
                      random_form="scalars"):
    """plonk::create_proof as a patched halo2_proofs would run it over the step API: transcript + RNG here, O(n) work there"""
    p = pkg()
    opts = opts or plonk.ProverOptions()
    cs = gpk.cs
    n, bf, A, L, Pn = cs.n, gpk.blinding_factors, cs.num_advice, gpk.num_lookups, gpk.num_sets
    if transcript_kind == "evm":
        tr = plonk.EvmTranscriptWrite()
    elif transcript_kind == "poseidon":
        from oracle.poseidon import PoseidonTranscriptWrite
        tr = PoseidonTranscriptWrite(opts.point_format)
    else:
        tr = plonk.TranscriptWrite(transcript_kind, opts.point_format)
    draw = rng.fr_random

    def write_points(arr):
        for pt in orc.g1_to_ints(arr):
            tr.write_point(pt)
    tr.common_scalar(transcript_repr)
    for col in instances:
        for v in col:
            tr.common_scalar(v)
    tails = None
    if opts.advice_blinding == "pse":
        tails = M([draw() for _ in range(A * (bf + 1))])
    if opts.blind_draws:
        for _ in range(A):
            draw()
    sess = p.ProverSession(gpk, advice, [M(c) for c in instances], advice_tails=tails)
    try:
        write_points(sess.advice_commitments)
        theta = tr.squeeze_challenge()
        t = []
        for _ in range(L):
            t += [draw() for _ in range(2 * (bf + 1))]
            if opts.blind_draws:
                draw(); draw()
        write_points(sess.lookups(M([theta]), M(t) if t else None, opts.lookup_fill))
        beta = tr.squeeze_challenge()
        gamma = tr.squeeze_challenge()
        t = []
        for _ in range(Pn + L):
            t += [draw() for _ in range(bf)]
            if opts.blind_draws:
                draw()
        write_points(sess.products(M([beta]), M([gamma]), M(t) if t else None))
        # vanishing argument: the caller describes the random polynomial in whichever form suits its RNG
        if opts.random_poly == "chunked":
            T = opts.random_poly_threads
            chunk = n // T
            seeds = [rng.fill_bytes(32) for _ in range(T + (1 if n % T else 0))]
            spec = p.RandomPolySpec(seeds=seeds, chunk_len=chunk)
        elif random_form == "keystream":
            assert isinstance(rng, orc.ChaCha20Rng)
            spec = p.RandomPolySpec(seed=rng.seed, rng="chacha20" if rng.rounds == 20 else "std", first_word=rng.word)
            rng.word += 16 * n
        else:
            spec = p.RandomPolySpec(scalars=rng.fr_random_bulk(n) if hasattr(rng, "fr_random_bulk") else M([draw() for _ in range(n)]))
        if opts.blind_draws:
            draw()
        write_points(sess.vanishing(spec))
        y = tr.squeeze_challenge()
        if opts.blind_draws:
            for _ in range(gpk.degree - 1):
                draw()
        write_points(sess.quotient(M([y])))
        x = tr.squeeze_challenge()
        for e in orc.fr_to_ints(sess.evals(M([x]))):
            tr.write_scalar(e)
        if multiopen == "shplonk":
            yy = tr.squeeze_challenge()
            v = tr.squeeze_challenge()
            write_points(sess.open_shplonk_h(M([yy]), M([v])))
            u = tr.squeeze_challenge()
            write_points(sess.open_shplonk_w(M([u])))
        else:
            v = tr.squeeze_challenge()
            write_points(sess.open_gwc(M([v])))
    finally:
        sess.end()
    return bytes(tr.proof)


@pytest.fixture(scope="module")
def circuit_k7():
    circ = pkg().synth.make_base_circuit(7, 2, seed=4)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    return circ, opk, advice, params, gpk


@pytest.mark.parametrize("transcript,multiopen", [("blake2b", "shplonk"), ("keccak", "gwc"), ("evm", "shplonk"), ("poseidon", "shplonk"),
                                                  ("poseidon", "gwc"), ("blake2b", "gwc")])
def test_step_api_with_host_transcript_matches_oracle(circuit_k7, transcript, multiopen):
    circ, opk, advice, params, gpk = circuit_k7
    seed = pyref.seed_from_u64(77)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20), transcript, multiopen)
    got = host_create_proof(gpk, opk.transcript_repr, np.concatenate(advice), circ.instances, pyref.ChaChaRng(seed, 20), transcript, multiopen)
    assert got == want
    # and it is what the library's own driver (zkc_prove: same rounds, built-in transcript + RNG) emits
    assert got == pkg().create_proof(gpk, np.concatenate(advice), [M(c) for c in circ.instances], seed, transcript, multiopen)


@pytest.mark.parametrize("opts", [dict(advice_blinding="pse"), dict(blind_draws=True), dict(point_format=1), dict(lookup_fill="axiom"),
                                  dict(random_poly="chunked", random_poly_threads=4),
                                  dict(random_poly="chunked", random_poly_threads=3, blind_draws=True, advice_blinding="pse"),
                                  dict(random_poly="chunked", random_poly_threads=1, blind_draws=True)])
def test_step_api_open_switches(circuit_k7, opts):
    circ, opk, advice, params, gpk = circuit_k7
    seed = pyref.seed_from_u64(78)
    o = plonk.ProverOptions(**opts)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20), opts=o)
    assert host_create_proof(gpk, opk.transcript_repr, np.concatenate(advice), circ.instances, pyref.ChaChaRng(seed, 20), opts=o) == want


@pytest.mark.parametrize("rounds", [20, 12])
def test_step_api_random_polynomial_forms(circuit_k7, rounds):
    """the random polynomial as n caller-drawn scalars, or as a keystream position of the caller's seeded RNG (produced on
    the device): same proof; StdRng = ChaCha12 (SURVEY OPEN-6)"""
    circ, opk, advice, params, gpk = circuit_k7
    seed = pyref.seed_from_u64(79)
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, rounds))
    adv = np.concatenate(advice)
    assert host_create_proof(gpk, opk.transcript_repr, adv, circ.instances, orc.ChaCha20Rng(seed, rounds), random_form="scalars") == want
    assert host_create_proof(gpk, opk.transcript_repr, adv, circ.instances, orc.ChaCha20Rng(seed, rounds), random_form="keystream") == want
    assert pkg().create_proof(gpk, adv, [M(c) for c in circ.instances], seed, rng="chacha20" if rounds == 20 else "std") == want


def test_step_api_order_and_errors(circuit_k7):
    """steps out of order are BAD_ARG and poison nothing they should not; one session per ctx; InvalidInstances on a wrong
    number of instance columns; a failed step leaves only zkc_prove_end valid"""
    p = pkg()
    circ, opk, advice, params, gpk = circuit_k7
    adv = np.concatenate(advice)
    inst = [M(c) for c in circ.instances]
    with pytest.raises(p.ZkcError) as e:
        p.ProverSession(gpk, adv, [])
    assert e.value.code == 10
    with pytest.raises(p.ZkcError) as e:
        p.create_proof(gpk, adv, inst + [M([1])], pyref.seed_from_u64(1))
    assert e.value.code == 10
    sess = p.ProverSession(gpk, adv, inst)
    with pytest.raises(p.ZkcError) as e:
        p.ProverSession(gpk, adv, inst)          # second session on the same ctx
    assert e.value.code == 1
    with pytest.raises(p.ZkcError) as e:
        sess.quotient(M([5]))                    # lookups / products / vanishing come first
    assert e.value.code == 1
    bf = gpk.blinding_factors
    sess.lookups(M([3]), M(range(1, 2 * (bf + 1) + 1)))
    sess.products(M([4]), M([5]), M(range(1, (gpk.num_sets + gpk.num_lookups) * bf + 1)))
    with pytest.raises(p.ZkcError):
        sess.vanishing(None)                     # no random polynomial given anywhere: the step fails ...
    with pytest.raises(p.ZkcError) as e:
        sess.vanishing(p.RandomPolySpec(scalars=M(range(1, circ.cs.n + 1))))
    assert e.value.code == 1                     # ... and the session is dead
    sess.end()
    # the ctx is usable again
    seed = pyref.seed_from_u64(2)
    assert p.create_proof(gpk, adv, inst, seed) == plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20))
