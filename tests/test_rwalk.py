"""
N2: the lock-step batched random walk that stands in for dynesty's per-chain ``evolve_point``
(bajes/inf/sampler/dynesty.py:503-614).  The golden file was produced by the UNMODIFIED ``sample_rwalk`` (oracle/
make_golden.py rwalk); a chain is a deterministic function of its random stream, so everything here is bit-exact.
"""
import numpy as np
import pytest

from bajes_b200.pool import BatchedPool
from bajes_b200.rwalk import LockstepRandomWalk, estimate_nmcmc, reflect, unitcheck
from oracle import bajes_port as port
from oracle import rwalk_cases as rc
from oracle import shims

from conftest import load_golden


def _row(res):
    u, v, logl, ncall, blob = res
    return np.concatenate([u, v, [logl, ncall, blob["accept"], blob["reject"], blob["fail"], blob["scale"]]])


@pytest.mark.parametrize("group", list(rc.GROUPS))
def test_oracle_port_reproduces_the_reference_chains(group):
    gold = load_golden("rwalk.npz")[group]
    g = rc.GROUPS[group]
    args, seeds = rc.make_args(group)
    got = np.array([_row(port.sample_rwalk_port(a, g["maxmcmc"], g["walks"], g["nact"], np.random.RandomState(s)))
                    for a, s in zip(args, seeds)])
    assert np.array_equal(got, gold)


@pytest.mark.parametrize("group", list(rc.GROUPS))
def test_lockstep_chains_equal_the_reference_chains(group):
    """All chains of a group advanced together, one batched likelihood call per proposal round."""
    gold = load_golden("rwalk.npz")[group]
    g = rc.GROUPS[group]
    args, seeds = rc.make_args(group)
    walk = LockstepRandomWalk(g["maxmcmc"], walks=g["walks"], nact=g["nact"], loglike_batch=rc.loglike_batch,
                              prior_transform_batch=rc.ptform_batch, draws="chain")
    got = np.array([_row(r) for r in walk.propose_batch(args, seeds=seeds)])
    assert np.array_equal(got, gold)
    # far fewer likelihood CALLS than evaluations: that is the point
    assert walk.n_evaluations >= int(gold[:, NCALL].sum())
    assert walk.n_batched_calls <= int(gold[:, NCALL].max()) + int(gold[:, NCALL + 3].max()) + 2
    # a single chain through the per-point entry gives the same answer
    one = walk.propose_batch(args[:1], seeds=seeds[:1])[0]
    assert np.array_equal(_row(one), gold[0])


NCALL = 2 * rc.NDIM + 1


def test_vector_draws_keep_the_contract():
    g = rc.GROUPS["wrapped"]
    args, _ = rc.make_args("wrapped")
    walk = LockstepRandomWalk(g["maxmcmc"], walks=g["walks"], nact=g["nact"], loglike_batch=rc.loglike_batch,
                              prior_transform_batch=rc.ptform_batch, seed=5)
    res = walk.propose_batch(args)
    assert len(res) == len(args)
    for (u, v, logl, ncall, blob), a in zip(res, args):
        assert u.shape == v.shape == (rc.NDIM,)
        assert np.all((u > 0) & (u < 1))
        assert np.array_equal(v, rc.ptform(u)) and logl == rc.loglike(v)
        assert logl > a[1]                                   # dynesty.py:574
        assert ncall == blob["accept"] + blob["reject"] and blob["accept"] >= 2
        assert set(blob) == {"accept", "reject", "fail", "scale"} and blob["scale"] == a[3]
    assert walk.propose_batch([]) == []


def test_pool_routes_evolve_point_to_the_lockstep_walk():
    g = rc.GROUPS["hard_edges"]
    args, seeds = rc.make_args("hard_edges")

    class Post:                                              # what BatchedPool and the walk see of a Posterior
        log_like_batch = staticmethod(rc.loglike_batch)
        prior_transform = staticmethod(rc.ptform)

    walk = LockstepRandomWalk(g["maxmcmc"], walks=g["walks"], nact=g["nact"], posterior=Post(), seed=1)
    pool = BatchedPool(Post(), processes=4)
    res = pool.map(walk.propose, args)                       # dynesty: self.M(evolve_point, args)
    assert len(res) == len(args) and walk.n_batched_calls < walk.n_evaluations / 4
    above = 0
    for (u, v, logl, ncall, blob), a in zip(res, args):
        assert np.array_equal(v, rc.ptform(u)) and logl == rc.loglike(v)
        above += logl > a[1]              # a chain that never decorrelated returns a prior draw (dynesty.py:603-607)
    assert above >= len(args) // 2


def test_helpers_match_the_reference_definitions():
    # dynesty.py:45-49
    x = np.array([-0.3, 0.2, 1.4, 2.3, -1.2])
    assert np.allclose(reflect(x), [0.3, 0.2, 0.6, 0.3, 0.8])
    # dynesty.py:27-42
    U = np.array([[0.1, 0.9, 0.5], [0.1, 1.2, 0.5], [0.0, 0.5, 0.5], [1.2, 0.5, -0.4]])
    assert unitcheck(U).tolist() == [True, False, False, False]
    assert unitcheck(U, np.array([False, False, False])).tolist() == [True, False, False, False]
    assert unitcheck(U, np.array([True, False, True])).tolist() == [True, True, False, False]
    assert unitcheck(U, np.array([False, True, False])).tolist() == [True, False, True, True]
    # inf/utils.py:140-161
    for ratio in (0.0, 0.01, 0.2, 0.5, 1.0):
        for old, mx in ((25, 400), (100, 4096), (10, 60)):
            assert int(estimate_nmcmc(ratio, old, mx)) == port._nmcmc_port(ratio, old, mx)


@pytest.mark.needs_reference
def test_port_vs_reference_live_on_fresh_seeds():
    if not shims.reference_available():
        pytest.skip("reference tree not present")
    import logging
    import sys
    import types
    shims.import_reference()
    sys.modules.setdefault("dynesty", types.SimpleNamespace(__version__="1.2.3"))
    from bajes.inf.sampler.dynesty import BajesDynestyProposal
    logging.disable(logging.WARNING)
    try:
        for group in ("wrapped", "hard_edges"):
            g = rc.GROUPS[group]
            args, _ = rc.make_args(group)
            prop = BajesDynestyProposal(g["maxmcmc"], walks=g["walks"], nact=g["nact"])
            for c, a in enumerate(args[:6]):
                np.random.seed(9000 + c)
                ref = _row(prop.propose(a))
                got = _row(port.sample_rwalk_port(a, g["maxmcmc"], g["walks"], g["nact"], np.random.RandomState(9000 + c)))
                assert np.array_equal(ref, got)
    finally:
        logging.disable(logging.NOTSET)


@pytest.mark.gpu
def test_lockstep_walk_on_the_gpu_equals_the_sequential_walk():
    """The same chains, once advanced in lock-step through the batched GPU likelihood and once one proposal at a time
    through the per-point GPU likelihood by the oracle's sequential restatement: identical points, counts and logL
    (a walker's logL does not depend on the batch it is evaluated in)."""
    from bajes_b200 import GWLikelihood, Posterior, synthetic as syn
    from bajes_b200.ensemble import BoxPrior
    cfg = syn.CONFIGS["C1"]
    cl = syn.product_classes()
    like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"])
    X, names = syn.draw_walkers(cfg, 256, 12, "narrow")
    lo, hi = X.min(axis=0), X.max(axis=0)
    periodic = [1 if n in ("ra", "psi", "phi_ref") else 0 for n in names]
    prior = BoxPrior(names, np.stack([lo, hi], axis=1), syn.constants(cfg), periodic=periodic)
    post = Posterior(like, prior)
    n = len(names)
    kwargs = dict(periodic=[i for i, p in enumerate(periodic) if p], reflective=[i for i, p in enumerate(periodic) if not p],
                  nonbounded=np.zeros(n, dtype=bool))
    rng = np.random.default_rng(3)
    U0 = rng.uniform(0.3, 0.7, size=(12, n))
    L0 = post.log_like_batch(post.prior_transform_batch(U0))
    args = [(U0[c], float(L0[c]) - 3.0, 0.05 * np.eye(n), 1.0, post.prior_transform, post.log_like, None, dict(kwargs))
            for c in range(len(U0))]
    seeds = list(range(40, 40 + len(U0)))
    walk = LockstepRandomWalk(80, walks=12, nact=3.0, posterior=post, draws="chain")
    got = walk.propose_batch(args, seeds=seeds)
    assert walk.n_batched_calls < walk.n_evaluations / 6
    for a, s, g in zip(args, seeds, got):
        ref = port.sample_rwalk_port(a, 80, 12, 3.0, np.random.RandomState(s))
        assert np.array_equal(g[0], ref[0]) and np.array_equal(g[1], ref[1])
        assert g[2] == ref[2] and g[3] == ref[3] and g[4] == ref[4]
