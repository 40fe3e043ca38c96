"""
Build-container tests (need the read-only reference tree): the oracle port against the UNMODIFIED reference run live
under oracle/shims.py, and a freshness check of the committed golden vectors.  Skipped where /root/reference does not
exist (the GPU box).
"""
import numpy as np
import pytest

from bajes_b200 import synthetic as syn
from oracle import harness

from conftest import load_golden

pytestmark = pytest.mark.needs_reference


@pytest.mark.parametrize("approx", ["TaylorF2_3.5PN", "TaylorF2_5.5PN", "TaylorF2_5.5PN_7.5PNTides",
                                    "TaylorF2_5.5PN_3.5PNQM_7.5PNTides", "TaylorF2_5.5PN_7.5PNTides2020"])
@pytest.mark.parametrize("marg", [False, True])
def test_port_matches_live_reference(approx, marg):
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=4.0, seed=777)
    ref = harness.reference_likelihood(cfg, approx=approx, marg_phi_ref=marg)
    gm = ref.dets["H1"].gmst_reference
    port = harness.port_likelihood(cfg, approx=approx, marg_phi_ref=marg, gmst_reference=gm, sday=ref.dets["H1"].sday)
    xa, names = syn.draw_walkers(cfg, 12, 5, "narrow")
    xb, _ = syn.draw_walkers(cfg, 12, 6, "wide")
    X = np.concatenate([xa, xb])
    X[:, names.index("lambda1")][::3] = 0.0
    X[:, names.index("lambda2")][::3] = 0.0      # both zero: the tidal branch is skipped (taylorf2.py:394)
    const = syn.constants(cfg)
    a = harness.eval_batch(ref, X, names, const)
    b = harness.eval_batch(port, X, names, const)
    assert np.all(np.abs(a - b) <= 1e-3 * harness.tolerance(a)), np.max(np.abs(a - b) / harness.tolerance(a))


def test_reference_time_marginalised_branch_matches_port():
    from bajes.pipe.log_like import GWLikelihood
    from oracle import bajes_port as port
    cfg = syn.CONFIGS["C1"]
    rc = harness.reference_classes()
    net = syn.build_network(cfg, rc)
    ref = GWLikelihood(net[0], net[1], net[2], net[3], net[4], cfg["srate"], cfg["seglen"], cfg["approx"],
                       marg_time_shift=True, marg_phi_ref=True)
    pc = harness.port_classes(gmst_reference=ref.dets["H1"].gmst_reference, sday=ref.dets["H1"].sday)
    pnet = syn.build_network(cfg, pc)
    mine = port.GWLikelihoodPort(pnet[0], pnet[1], pnet[2], pnet[3], pnet[4], cfg["srate"], cfg["seglen"],
                                 cfg["approx"], marg_time_shift=True, marg_phi_ref=True)
    p = syn.injection_params(cfg)
    assert abs(ref.log_like(dict(p)) - mine.log_like(dict(p))) < 1e-6


def test_golden_c1_is_reproducible_from_the_reference():
    """The committed C1 vectors are what the reference produces today (guards against a stale fixture)."""
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    ref = harness.reference_likelihood(cfg)
    names = [str(n) for n in gold["names"]]
    got = harness.eval_batch(ref, gold["X"][:16], names, syn.constants(cfg))
    np.testing.assert_allclose(got, gold["logl_marg0"][:16], rtol=0, atol=1e-9)
    np.testing.assert_allclose([ref.dets[i].gmst_reference for i in cfg["ifos"]], gold["gmst_reference"])


def test_detector_mirror_geometry_matches_the_reference_methods():
    """The scalar geometry helpers the post-processing scripts call on ``Detector`` (detector.py:246-392): same numbers
    from the product's mirror class and from the unmodified reference class, fed the same GMST reference."""
    from bajes_b200 import synthetic as syn
    from bajes_b200.detector import Detector
    cl = harness.reference_classes()
    rng = np.random.default_rng(4)
    for ifo in ("H1", "L1", "V1"):
        ref = cl.Detector(ifo, syn.T_GPS)
        mine = Detector(ifo, syn.T_GPS)
        mine.gmst_reference, mine.sday = ref.gmst_reference, ref.sday
        other = cl.Detector("V1" if ifo != "V1" else "H1", syn.T_GPS)
        for _ in range(20):
            ra, dec, psi = rng.uniform(0, 2 * np.pi), np.arcsin(rng.uniform(-1, 1)), rng.uniform(0, np.pi)
            t = syn.T_GPS + rng.uniform(-100, 100)
            assert mine.gmst_estimate(t) == ref.gmst_estimate(t)
            assert np.isclose(mine.time_delay_from_earth_center(ra, dec, t), ref.time_delay_from_earth_center(ra, dec, t),
                              rtol=1e-13, atol=1e-18)
            assert np.isclose(mine.time_delay_from_detector(other, ra, dec, t),
                              ref.time_delay_from_detector(other, ra, dec, t), rtol=1e-13, atol=1e-18)
            loc = rng.uniform(-6e6, 6e6, 3)
            assert np.isclose(mine.time_delay_from_location(loc, ra, dec, t), ref.time_delay_from_location(loc, ra, dec, t),
                              rtol=1e-13, atol=1e-18)
            assert np.allclose(mine.antenna_pattern(ra, dec, psi, t), ref.antenna_pattern(ra, dec, psi, t), rtol=1e-12,
                               atol=1e-15)
            assert np.allclose(mine.optimal_orientation(t), ref.optimal_orientation(t), rtol=1e-14)
        assert np.isclose(mine.light_travel_time_to_detector(other), ref.light_travel_time_to_detector(other), rtol=1e-14)
