"""GPU: keygen behind the C ABI — zkc_keygen_pk (fixed columns + copy constraints -> resident proving
key; what snark-verifier-sdk's gen_pk does, /root/reference/src/helpers.rs:213,265) and ProvingKey files (zkc_pk_write /
zkc_pk_read; /root/reference/src/bin/cli.rs:247,312) against the oracle's keygen and prover."""
import numpy as np
import pytest

from oracle import orc, plonk
from tests import pyref
from tests.circuits import oracle_setup, oracle_srs
from tests.util import gpu_ctx, pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup_k8():
    p = pkg()
    circ = p.synth.make_base_circuit(8, 3, seed=21)
    opk, advice = oracle_setup(circ)
    g, gl = oracle_srs(8)
    params = p.ParamsKZG(8, g=g, g_lagrange=gl, ctx=gpu_ctx())
    fixed = np.concatenate(opk.fixed_values)
    tr = orc.fr_from_ints([opk.transcript_repr])
    return circ, opk, advice, params, fixed, tr


def test_keygen_pk_from_copy_constraints(setup_k8):
    p = pkg()
    circ, opk, advice, params, fixed, tr = setup_k8
    pk = p.ProvingKey.keygen(params, circ.cs, fixed, circ.copies, tr)
    f, s = pk.commitments()
    assert orc.g1_to_ints(f) == opk.fixed_commitments and orc.g1_to_ints(s) == opk.sigma_commitments
    assert np.array_equal(pk.sigma(), np.concatenate(opk.sigma_values))            # pk.permutation.permutations, byte for byte
    seed = pyref.seed_from_u64(3)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    assert p.create_proof(pk, np.concatenate(advice), inst, seed) == plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20))
    with pytest.raises(p.ZkcError) as e:        # a copy that leaves the permutation columns: plonk::Error::BoundsFailure
        p.ProvingKey.keygen(params, circ.cs, fixed, list(circ.copies) + [(len(circ.cs.permutation), 0, 0, 0)], tr)
    assert e.value.code == 12


@pytest.mark.parametrize("be", [True, False])
def test_pk_file_round_trip(setup_k8, be):
    p = pkg()
    circ, opk, advice, params, fixed, tr = setup_k8
    pk = p.ProvingKey.keygen(params, circ.cs, fixed, circ.copies, tr)
    n = circ.cs.n
    selectors = np.packbits(np.array([[1 if v else 0 for v in circ.fixed[i]] for i in range(3)], dtype=np.uint8), axis=1, bitorder="little")
    data = pk.write(selectors, be=be)
    L = p.api.pk_file_layout(pk.k, pk.extended_k, circ.cs.num_fixed, len(circ.cs.permutation), 3)
    assert len(data) == L.total
    # the sections hold what the oracle's keygen holds (Montgomery limbs, as RawBytes writes them)
    col = lambda off, c, ln: np.frombuffer(data, dtype=np.uint8, count=32 * ln, offset=off + 4 + c * (4 + 32 * ln) + 4).view(np.uint64).reshape(ln, 4)
    en = 1 << pk.extended_k
    assert np.array_equal(col(L.fixed_values_off, 1, n), opk.fixed_values[1])
    assert np.array_equal(col(L.fixed_polys_off, 2, n), opk.fixed_polys[2])
    assert np.array_equal(col(L.fixed_cosets_off, 0, en), opk.fixed_cosets[0])
    assert np.array_equal(col(L.perm_values_off, 3, n), opk.sigma_values[3])
    assert np.array_equal(col(L.perm_polys_off, 0, n), opk.sigma_polys[0])
    assert np.array_equal(col(L.perm_cosets_off, 5, en), opk.sigma_cosets[5])
    ext = lambda off: np.frombuffer(data, dtype=np.uint8, count=32 * en, offset=off + 4).view(np.uint64).reshape(en, 4)
    assert np.array_equal(ext(L.l0_off), opk.l0) and np.array_equal(ext(L.l_last_off), opk.l_last) and np.array_equal(ext(L.l_active_row_off), opk.l_active_row)
    assert data[L.selectors_off:L.l0_off] == selectors.tobytes()
    # read it back: same key, same proofs — through both SerdeFormats
    seed = pyref.seed_from_u64(4)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    want = plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20))
    for fmt in ("unchecked", "raw"):
        pk2 = p.ProvingKey.read(params, circ.cs, data, tr, num_selectors=3, file_format=fmt)
        f, s = pk2.commitments()
        assert orc.g1_to_ints(f) == opk.fixed_commitments and orc.g1_to_ints(s) == opk.sigma_commitments
        assert p.create_proof(pk2, np.concatenate(advice), inst, seed) == want
    # RawBytes is checked: a tampered column no longer matches the vk commitments; a non-canonical scalar is rejected;
    # RawBytesUnchecked trusts the file (as upstream)
    bad = bytearray(data)
    bad[L.fixed_values_off + 8 + 5 * 32] ^= 1
    with pytest.raises(p.ZkcError):
        p.ProvingKey.read(params, circ.cs, bytes(bad), tr, num_selectors=3, file_format="raw")
    p.ProvingKey.read(params, circ.cs, bytes(bad), tr, num_selectors=3, file_format="unchecked")
    bad = bytearray(data)
    bad[L.perm_values_off + 8:L.perm_values_off + 8 + 32] = b"\xff" * 32
    with pytest.raises(p.ZkcError):
        p.ProvingKey.read(params, circ.cs, bytes(bad), tr, num_selectors=3, file_format="raw")
    with pytest.raises(p.ZkcError):           # wrong shape (selector count) is an error in either format
        p.ProvingKey.read(params, circ.cs, data, tr, num_selectors=2)
