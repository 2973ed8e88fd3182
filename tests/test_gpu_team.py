"""Team proving (one create_proof over several GPUs) — the sharding arithmetic on ONE GPU.

`zkc_team_emulate(W)` makes the context run the W shards of every partitioned step (point-range MSM slices,
column blocks of the transforms, extended-row blocks of h(X)) one after the other with the collectives elided,
so the bytes must equal the unsharded proof and the oracle's.  The NCCL path itself is covered by
tools/team_check.py under torchrun on a multi-GPU box."""
import numpy as np
import pytest

from oracle import orc, plonk
from tests import pyref
from tests.circuits import oracle_setup
from tests.test_gpu_prover import gpu_setup
from tests.util import gpu_ctx, pkg

pytestmark = pytest.mark.gpu


@pytest.fixture
def team_ctx():
    ctx = gpu_ctx()
    yield ctx
    ctx.team_emulate(1)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("maker,args", [("make_base_circuit", (8, 3)), ("make_multi_lookup_circuit", (7,)), ("make_sha_bit_circuit", (9, 48, 3))])
def test_emulated_team_proof_equals_single_gpu_and_oracle(team_ctx, world, maker, args):
    p = pkg()
    kw = dict(blocks=4, seed=2) if maker == "make_sha_bit_circuit" else dict(seed=5)
    circ = getattr(p.synth, maker)(*args, **kw)
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    seed = pyref.seed_from_u64(world)
    for transcript, multiopen in (("blake2b", "shplonk"), ("keccak", "gwc")):
        single = p.create_proof(gpk, np.concatenate(advice), inst, seed, transcript, multiopen)
        team_ctx.team_emulate(world)
        assert team_ctx.team_info() == (0, world, True)
        team = p.create_proof(gpk, np.concatenate(advice), inst, seed, transcript, multiopen)
        team_ctx.team_emulate(1)
        assert team == single
        assert team == plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20), transcript, multiopen)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_emulated_team_commitments_by_column(team_ctx, world):
    """SURVEY 8e row 5 (per-column commitments dealt to the GPUs): with `team_commit_by_column` a batch of at least `world`
    commitments is split by column (rank r commits its columns whole) instead of by point range — same points, same proof;
    batches with fewer columns than ranks keep the point-range split."""
    p = pkg()
    circ = p.synth.make_sha_bit_circuit(9, 48, 3, blocks=4, seed=2)      # 51 advice columns: 17 / 17 / 17 and 13 / 13 / 13 / 12
    opk, advice = oracle_setup(circ)
    params, gpk = gpu_setup(circ, opk)
    inst = [orc.fr_from_ints(c) for c in circ.instances]
    seed = pyref.seed_from_u64(40 + world)
    single = p.create_proof(gpk, np.concatenate(advice), inst, seed)
    from tests.util import to_dev
    cols = to_dev(np.concatenate(advice[:7]))
    want = params.commit_dev(cols, 1 << 9, 7, basis=1)
    team_ctx.team_emulate(world)
    team_ctx.set_tunable("team_commit_by_column", 1)
    try:
        assert p.create_proof(gpk, np.concatenate(advice), inst, seed) == single
        assert np.array_equal(params.commit_dev(cols, 1 << 9, 7, basis=1), want)
    finally:
        team_ctx.set_tunable("team_commit_by_column", 0)
        team_ctx.team_emulate(1)
    assert single == plonk.create_proof(opk, advice, circ.instances, pyref.ChaChaRng(seed, 20))


def test_emulated_team_keygen_and_commit(team_ctx):
    """pk_load (vk commitments) and ParamsKZG.commit under a team context: point-range shards sum to the same points"""
    p = pkg()
    circ = p.synth.make_base_circuit(9, 2, seed=9)
    opk, advice = oracle_setup(circ)
    team_ctx.team_emulate(4)
    params, gpk = gpu_setup(circ, opk)
    f, s = gpk.commitments()
    assert orc.g1_to_ints(f) == opk.fixed_commitments and orc.g1_to_ints(s) == opk.sigma_commitments
    team_ctx.team_emulate(1)


def test_emulated_team_full_size(team_ctx):
    """RSA k=17 shape, 8 shards: same bytes as the unsharded proof"""
    p = pkg()
    w = p.workload.build(team_ctx, 17, 3, seed=11)
    seed = pyref.seed_from_u64(17)
    single = p.create_proof(w.pk, w.advice_dev, w.instances, seed)
    team_ctx.team_emulate(8)
    assert p.create_proof(w.pk, w.advice_dev, w.instances, seed) == single


def test_aggregation_shape_k20_verifies_and_team_matches(team_ctx):
    """BASELINE config-5 shape (BaseConfig, 17 gate columns + lookup, 20 permutation columns) reduced to k=20 (the k=22
    build needs ~65 GB): the proof verifies under the independent verifier, host-buffer (staged upload) == device-resident,
    and the 4-shard team proof has the same bytes."""
    from tests.test_gpu_prover import _verify_workload
    p = pkg()
    w = p.workload.build(team_ctx, 20, 17, seed=100, shape="base_fast")
    seed = pyref.seed_from_u64(20)
    proof = p.create_proof(w.pk, w.advice_dev, w.instances, seed)
    assert _verify_workload(w, proof)
    assert p.create_proof(w.pk, w.advice_host, w.instances, seed) == proof      # 604 MB witness: staged upload path
    team_ctx.team_emulate(4)
    assert p.create_proof(w.pk, w.advice_dev, w.instances, seed) == proof
    assert p.create_proof(w.pk, w.advice_host, w.instances, seed) == proof      # team: 1/W upload shares
