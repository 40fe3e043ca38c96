"""
GPU edge cases (``-m gpu``): ragged and empty batches, odd / tiny frequency bands, 1 and 2 detectors, alternative
parametrisations (mtot, iota, spherical and in-plane spins, constants vs sampled), pickling, the pool adapter.
Every value is compared with the oracle port evaluated live on the host.
"""
import pickle

import numpy as np
import pytest

from bajes_b200 import GWLikelihood, Posterior, synthetic as syn
from bajes_b200.ensemble import BoxPrior
from bajes_b200.pool import BatchedPool
from oracle import harness

from conftest import product_network_from_port

pytestmark = pytest.mark.gpu


def _pair(cfg, **kw):
    """(GPU likelihood, oracle-port likelihood) on identical data."""
    cl = harness.port_classes()
    net = syn.build_network(cfg, cl)
    port = harness.port_likelihood(cfg, net=net, marg_phi_ref=kw.get("marg_phi_ref", False))
    ifos, datas, dets, noises, f = product_network_from_port(cfg, net)
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], **kw)
    return like, port


def _agree(got, want, what):
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), what
    err = np.abs(got[fin] - want[fin])
    assert np.all(err <= harness.tolerance(want[fin])), "%s: %g x tol" % (what, np.max(err / harness.tolerance(want[fin])))


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 300])
def test_ragged_and_empty_batches(n):
    cfg = syn.CONFIGS["C1"]
    like, port = _pair(cfg)
    X, names = syn.draw_walkers(cfg, max(n, 1), 3, "wide")
    X = X[:n]
    const = syn.constants(cfg)
    got = like.log_like_batch(X, names, const)
    assert got.shape == (n,)
    if n:
        _agree(got, harness.eval_batch(port, X, names, const), "batch of %d" % n)


@pytest.mark.parametrize("band", [(20.0, 21.0), (20.0, 36.0), (20.25, 100.0), (30.0, 511.75)])
def test_tiny_odd_and_even_bands(band):
    """5, 65, 320 and 1928 in-band bins: below one tile, one bin over a tile, odd and even counts."""
    cfg = dict(syn.CONFIGS["C1"])
    cfg.update(f_min=band[0], f_max=band[1])
    like, port = _pair(cfg, sincos="poly32")
    n_bins = int(round((band[1] - band[0]) * cfg["seglen"])) + 1
    assert like.engine.n_bins == n_bins
    X, names = syn.draw_walkers(cfg, 40, 4, "wide")
    const = syn.constants(cfg)
    _agree(like.log_like_batch(X, names, const), harness.eval_batch(port, X, names, const), "band %s" % (band,))


@pytest.mark.parametrize("ifos", [["V1"], ["L1", "K1"], ["H1", "L1", "V1"]])
def test_detector_networks(ifos):
    from bajes_b200 import noise as nz
    cfg = dict(syn.CONFIGS["C1"])
    cfg.update(ifos=ifos)
    if "K1" in ifos:
        # no KAGRA curve is bundled: reuse the aLIGO table for the test (both sides get the same PSD)
        nz._DESIGN_KEY["K1"] = "aLIGO"
    like, port = _pair(cfg)
    X, names = syn.draw_walkers(cfg, 48, 5, "wide")
    const = syn.constants(cfg)
    _agree(like.log_like_batch(X, names, const), harness.eval_batch(port, X, names, const), "ifos %s" % ifos)
    nz._DESIGN_KEY.pop("K1", None)


def test_alternative_parametrisations():
    """mtot instead of mchirp, iota instead of cos_iota, spherical spins, in-plane cartesian spins, and a mix of
    sampled columns and constants (waveform.py:334-377)."""
    cfg = syn.CONFIGS["C2_small"]
    like, port = _pair(cfg, sincos="poly32")
    base = syn.constants(cfg)
    rng = np.random.default_rng(8)
    n = 24
    cols = dict(mtot=rng.uniform(2.5, 3.2, n), q=rng.uniform(1.0, 1.8, n), s1=rng.uniform(0, 0.3, n),
                tilt1=rng.uniform(0, np.pi, n), phi_1l=rng.uniform(0, 2 * np.pi, n), s2=rng.uniform(0, 0.3, n),
                tilt2=rng.uniform(0, np.pi, n), phi_2l=rng.uniform(0, 2 * np.pi, n), lambda1=rng.uniform(0, 1000, n),
                iota=rng.uniform(0, np.pi, n), ra=rng.uniform(0, 6.28, n), dec=rng.uniform(-1.5, 1.5, n),
                psi=rng.uniform(0, 3.14, n), time_shift=rng.uniform(-0.01, 0.01, n), phi_ref=rng.uniform(0, 6.28, n))
    names = list(cols)
    X = np.stack([cols[k] for k in names], axis=1)
    const = {k: v for k, v in base.items() if k not in ("s1x", "s1y", "s2x", "s2y")}
    const.update(lambda2=250.0, distance=60.0)            # constants instead of sampled columns
    got = like.log_like_batch(X, names, const)
    _agree(got, harness.eval_batch(port, X, names, const), "mtot / iota / spherical spins")
    # in-plane cartesian components enter the 3.5PN spin-spin terms (taylorf2.py:199-205)
    names2 = ["mchirp", "q", "s1x", "s1y", "s1z", "s2x", "s2y", "s2z", "cos_iota", "ra", "dec", "psi", "time_shift",
              "phi_ref", "distance"]
    X2 = np.stack([rng.uniform(1.1, 1.3, n), rng.uniform(1, 2, n)] + [rng.uniform(-0.2, 0.2, n) for _ in range(6)] +
                  [rng.uniform(-1, 1, n), rng.uniform(0, 6.28, n), rng.uniform(-1.5, 1.5, n), rng.uniform(0, 3.14, n),
                   rng.uniform(-0.01, 0.01, n), rng.uniform(0, 6.28, n), rng.uniform(20, 80, n)], axis=1)
    const2 = {k: v for k, v in base.items() if k not in ("s1x", "s1y", "s2x", "s2y")}
    const2.update(lambda1=0.0, lambda2=0.0)                # both zero: the tidal branch is skipped (taylorf2.py:394)
    _agree(like.log_like_batch(X2, names2, const2), harness.eval_batch(port, X2, names2, const2), "cartesian spins")
    # per-point dict API with the same content
    p = {k: float(v) for k, v in zip(names2, X2[0])}
    p.update(const2)
    assert abs(like.log_like(dict(p)) - port.log_like(dict(p))) <= float(harness.tolerance(port.log_like(dict(p))))


def test_missing_required_parameter_raises():
    from bajes_b200.engine import EngineError
    cfg = syn.CONFIGS["C1"]
    like, _ = _pair(cfg)
    X, names = syn.draw_walkers(cfg, 4, 1, "narrow")
    const = syn.constants(cfg)
    keep = [i for i, n in enumerate(names) if n != "distance"]
    with pytest.raises(EngineError):
        like.log_like_batch(X[:, keep], [names[i] for i in keep], const)
    with pytest.raises(ValueError):
        like.log_like_batch(X[:, :5], names, const)


def test_pickle_round_trip_rebuilds_the_engine_lazily():
    cfg = syn.CONFIGS["C1"]
    like, _ = _pair(cfg, marg_phi_ref=True)
    X, names = syn.draw_walkers(cfg, 16, 2, "narrow")
    const = syn.constants(cfg)
    a = like.log_like_batch(X, names, const)
    clone = pickle.loads(pickle.dumps(like))
    assert clone._engine is None                       # device handles never travel
    assert np.array_equal(clone.log_like_batch(X, names, const), a)
    assert clone.dets["H1"]._owner is clone


def test_posterior_and_pool_adapter_on_the_gpu():
    """pool.map(posterior.log_post / log_likeprior / log_like) == the per-point Posterior, row by row."""
    cfg = syn.CONFIGS["C1"]
    like, port = _pair(cfg)
    X, names = syn.draw_walkers(cfg, 33, 6, "wide")
    lo, hi = X.min(axis=0), X.max(axis=0)
    prior = BoxPrior(names, np.stack([lo - 1e-9, hi + 1e-9], axis=1), syn.constants(cfg))
    Xq = X.copy()
    Xq[4, 0] = hi[0] + 1.0                              # one point outside the prior box
    post = Posterior(like, prior)
    ref_post = Posterior(port, prior)
    pool = BatchedPool(post, processes=4)
    lp = np.array(pool.map(post.log_post, list(Xq)))
    assert lp[4] == -np.inf and pool.n_batched_calls == 1
    want = np.array([ref_post.log_post(x) for x in Xq])
    keep = np.isfinite(want)
    assert np.array_equal(np.isfinite(lp), keep)
    assert np.all(np.abs(lp[keep] - want[keep]) <= harness.tolerance(want[keep]))
    lnl_lnp = pool.map(post.log_likeprior, list(Xq))
    assert lnl_lnp[4] == (0.0, -np.inf)                 # inf/likelihood.py:79-80
    ll = np.array(pool.map(post.log_like, list(Xq[:8])))
    assert np.all(np.abs(ll - np.array([ref_post.log_like(x) for x in Xq[:8]])) <= harness.tolerance(ll))


def test_frequency_split_partial_sums_recombine():
    """The frequency-axis split used for small batches on several GPUs, emulated on one GPU: partial sums of disjoint
    chunk ranges, added (the all-reduce), then the epilogue == the single-call result to rounding."""
    import torch
    cfg = syn.CONFIGS["C2_small"]
    ex = syn.extras_setup(cfg, 4, 3)
    cl = syn.product_classes()
    like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"], **ex)
    X, names = syn.draw_walkers(cfg, 50, 9, "narrow")
    Xe, ne = syn.extras_columns(cfg, 50, 10, 4, 3)
    X, names = np.ascontiguousarray(np.concatenate([X, Xe], axis=1)), names + ne
    const = syn.constants(cfg)
    whole = like.log_like_batch(X, names, const)
    dev = torch.device("cuda", 0)
    pm = like.param_map(names, const)
    xd = torch.from_numpy(X).to(dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    nch = like.engine.num_chunks()
    assert nch >= 3
    cuts = [0, nch // 3, nch // 3, 2 * nch // 3 + 1, nch]            # includes an empty range
    total = torch.zeros((2 * len(like.ifos), len(X)), dtype=torch.float64, device=dev)
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = torch.empty_like(total)
        like.engine.partial_device(xd.data_ptr(), len(X), X.shape[1], pm, a, b, part.data_ptr(), st)
        total += part
    out = torch.empty(len(X), dtype=torch.float64, device=dev)
    like.engine.finish_device(total.data_ptr(), len(X), out.data_ptr(), 0, st)
    torch.cuda.synchronize(dev)
    got = out.cpu().numpy()
    assert np.all(np.abs(got - whole) <= 1e-9 * (np.abs(whole) + 1.0)), np.max(np.abs(got - whole))
    # and the single-rank FrequencySplitEvaluator path
    from bajes_b200.dist import FrequencySplitEvaluator
    ev = FrequencySplitEvaluator(like, names, const)
    assert np.array_equal(ev.evaluate(X), whole) or np.all(np.abs(ev.evaluate(X) - whole) <= 1e-9 * (np.abs(whole) + 1))


# ------------------------------------------------------------------------------------------------------
# networks of more than three detectors (the reference loops over any number: log_like.py:57-96; ET1-3, K1, CE in
# detector.py:26-29): evaluated in passes of three
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("ifos", [["H1", "L1", "V1", "K1"], ["H1", "L1", "V1", "K1", "ET1", "ET2", "ET3"]])
def test_networks_of_more_than_three_detectors(ifos, marg):
    from types import SimpleNamespace
    from conftest import design_asd_any, product_network_from_port
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(ifos=list(ifos), seglen=8.0, seed=5105)
    pc = harness.port_classes()
    cl = SimpleNamespace(**{**vars(pc), "design_asd": design_asd_any})
    net = syn.build_network(cfg, cl)
    port_like = harness.port_likelihood(cfg, marg_phi_ref=marg, net=net)
    pifos, datas, dets, noises, f = product_network_from_port(cfg, net)
    like = GWLikelihood(pifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], marg_phi_ref=marg)
    assert len(like.engines) == -(-len(ifos) // 3)
    Xa, names = syn.draw_walkers(cfg, 48, 61, "narrow")
    Xb, _ = syn.draw_walkers(cfg, 16, 62, "wide")
    X = np.concatenate([Xa, Xb])
    X[-1, names.index("distance")] = np.inf
    const = syn.constants(cfg)
    with np.errstate(all="ignore"):
        ref = harness.eval_batch(port_like, X, names, const)
    got = like.log_like_batch(X, names, const)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and np.all(np.isneginf(got[~fin]))
    assert np.all(np.abs(got[fin] - ref[fin]) <= harness.tolerance(ref[fin]))
    p = {k: float(v) for k, v in zip(names, X[3])}
    p.update(const)
    assert like.log_like(dict(p)) == got[3]
    ip = like.inner_products(dict(p))
    assert sorted(ip) == sorted(ifos)
    h = like.project_fdwave(dict(p))
    assert sorted(h) == sorted(ifos) and len(h[ifos[-1]]) == len(h[ifos[0]]) == like.engines[0].n_bins


def test_batches_beyond_two_million_walkers_are_sliced():
    """walker tiles are a grid.y coordinate (<= 65 535 tiles of 64): the C ABI evaluates larger batches in slices of 2^21;
    every walker gets the bits it gets in a small call"""
    cfg = dict(syn.CONFIGS["C1"])
    cfg.update(f_max=128.0)                      # 433 bins: the batch, not the band, is what is large here
    like, _ = _pair(cfg)
    n = (1 << 21) + 4099
    X, names = syn.draw_walkers(cfg, 4099, 7, "wide")
    big = np.tile(X, (n // len(X) + 1, 1))[:n]
    got = like.log_like_batch(big, names, syn.constants(cfg))
    small = like.log_like_batch(X, names, syn.constants(cfg))
    assert got.shape == (n,)
    assert np.array_equal(got[:4099], small) and np.array_equal(got[-4099:], np.tile(small, 3)[(n - 4099) % 4099:][:4099])
