"""
TEST INFRASTRUCTURE.  Seeded toy problems for the dynesty random-walk state machine:
shared by ``oracle/make_golden.py rwalk`` (which runs the UNMODIFIED ``BajesDynestyProposal.sample_rwalk``,
bajes/inf/sampler/dynesty.py:503-614, on them) and by the tests of ``oracle.bajes_port.sample_rwalk_port`` and of the
product's lock-step ``bajes_b200.rwalk.LockstepRandomWalk``.
"""
from __future__ import annotations

import numpy as np

NDIM = 5
MU = np.array([0.7, -1.1, 0.3, 2.0, -0.4])
SIG = np.array([0.8, 1.3, 0.5, 1.0, 2.0])
LO, HI = -5.0, 5.0


def ptform_batch(U):
    return LO + (HI - LO) * np.asarray(U, dtype=np.float64)


def loglike_batch(V):
    """Fixed summation order, so a row's value does not depend on the batch it is evaluated in."""
    t = ((np.asarray(V, dtype=np.float64) - MU) / SIG) ** 2
    acc = np.zeros(len(t))
    for j in range(NDIM):
        acc = acc + t[:, j]
    return -0.5 * acc


def ptform(u):
    return ptform_batch(np.asarray(u)[None])[0]


def loglike(v):
    return float(loglike_batch(np.asarray(v)[None])[0])


# name -> (proposal settings, boundary kwargs, how loglstar is chosen)
GROUPS = {
    # bajes' own set-up: every parameter periodic or reflective, dynesty's nonbounded mask all False (dynesty.py:183-207)
    "wrapped": dict(maxmcmc=400, walks=25, nact=5.0, n=24, seed=100,
                    kwargs=dict(periodic=[0], reflective=[1, 2, 3, 4], nonbounded=np.zeros(NDIM, dtype=bool)),
                    star="start"),
    # hard edges on three parameters: proposals leave the cube (nfail branch, dynesty.py:559-568)
    "hard_edges": dict(maxmcmc=300, walks=20, nact=5.0, n=16, seed=200,
                       kwargs=dict(periodic=[0], reflective=[3], nonbounded=np.array([False, True, True, False, True])),
                       star="start", scale=2.5),
    # nothing is ever accepted: maxmcmc is hit and a prior draw is returned (dynesty.py:593-607)
    "never": dict(maxmcmc=60, walks=10, nact=5.0, n=6, seed=300,
                  kwargs=dict(periodic=[0], reflective=[1, 2, 3, 4], nonbounded=np.zeros(NDIM, dtype=bool)),
                  star="inf"),
    # everything is accepted
    "always": dict(maxmcmc=200, walks=15, nact=2.0, n=6, seed=400,
                   kwargs=dict(periodic=None, reflective=[0, 1, 2, 3, 4], nonbounded=np.zeros(NDIM, dtype=bool)),
                   star="-inf"),
}


def make_args(group):
    """The tuples dynesty (>= 1.2) hands to ``evolve_point``, plus the per-chain seeds."""
    g = GROUPS[group]
    rng = np.random.default_rng(g["seed"])
    args, seeds = [], []
    for c in range(g["n"]):
        u = np.clip(0.5 + (MU + 0.5 * SIG * rng.standard_normal(NDIM)) / (HI - LO), 0.02, 0.98)
        A = 0.08 * (np.eye(NDIM) + 0.3 * rng.standard_normal((NDIM, NDIM)))
        scale = float(g.get("scale", 1.0) * rng.uniform(0.5, 1.5))
        l0 = loglike(ptform(u))
        loglstar = {"start": l0 - rng.uniform(0.05, 2.0), "inf": 1e9, "-inf": -np.inf}[g["star"]]
        seed = g["seed"] * 1000 + c
        args.append((u, loglstar, A, scale, ptform, loglike, None, dict(g["kwargs"])))
        seeds.append(seed)
    return args, seeds
