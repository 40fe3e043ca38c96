"""
One test per reference quirk of SURVEY.md section 8(a) -- behaviours a faithful implementation must reproduce.  Each test
names the reference lines, shows that the oracle restatement HAS the quirk (a variant without it differs measurably) and
that the GPU path agrees with the oracle, i.e. has it too.

Quirks 9 (antenna pattern evaluated at t + delay, the delay itself at t), 11 (signs of the centre shift and of the
detector ramp) and 14 (ROQ spline outside its range -> -inf) are covered where the quantities are compared directly:
tests/test_host.py::test_detector_mirror_geometry_matches_the_reference_methods,
tests/test_gpu_parity.py::test_c1_projected_waveform_matches_oracle and ::test_roq_many_points_vs_port.
"""
import numpy as np
import pytest

from bajes_b200 import GWLikelihood, synthetic as syn
from oracle import bajes_port as port
from oracle import harness

from conftest import product_network_from_port

pytestmark = pytest.mark.gpu

NAMES = syn.WALKER_NAMES


def _pair(name="C2_small", approx=None, **over):
    cfg = dict(syn.CONFIGS[name])
    cfg.update(over)
    approx = approx or cfg["approx"]
    net = syn.build_network(cfg, harness.port_classes())
    ref = harness.port_likelihood(cfg, approx=approx, net=net)
    like = GWLikelihood(*product_network_from_port(cfg, net), cfg["srate"], cfg["seglen"], approx)
    return cfg, like, ref


def _point(cfg, **over):
    p = syn.injection_params(cfg)
    p.update(over)
    return p


def _close(a, b):
    return abs(a - b) <= harness.tolerance(np.array([b]))[0]


def test_quirk01_gt_literal_is_not_mtsun_si(monkeypatch):
    """taylorf2.py:6 vs bajes/__init__.py:68: v uses gt = 4.92549094830932e-6, NOT MTSUN_SI = 4.925491025543576e-06
    (1.6e-8 relative: ~3e-3 rad of phase over a long inspiral)."""
    cfg, like, ref = _pair()
    # at half the injected distance logL ~ 0 (2 <d|h> - 2 <h|h>): the tolerance is at its floor and the effect is plain
    p = _point(cfg, distance=0.5 * syn.INJECTIONS[cfg["source"]]["distance"])
    with_gt = ref.log_like(dict(p))
    assert _close(like.log_like(dict(p)), with_gt)
    monkeypatch.setattr(port, "GT", 4.925491025543576e-06)
    with_mtsun = ref.log_like(dict(p))
    assert abs(with_mtsun - with_gt) > 10 * harness.tolerance(np.array([with_gt]))[0]


def test_quirk02_amplitude_prefactor_literal():
    """taylorf2.py:379: pre = 3.6686934875530996e-19 as a literal; <h|h> = pre^2 (...) checked to 1e-12 against the oracle."""
    cfg, like, ref = _pair("C1")
    p = _point(cfg)
    assert port.PRE == 3.6686934875530996e-19 and port.EULER_GAMMA == 0.57721566490153286060
    got = like.inner_products(dict(p))
    hphc = ref.wave.compute_hphc(dict(p))
    for ifo in like.ifos:
        _, hh, _ = ref.dets[ifo].compute_inner_products(hphc, dict(p), "freq")
        assert abs(got[ifo][1] - np.real(hh)) <= 1e-11 * np.real(hh)


def test_quirk03_no_isco_cutoff():
    """taylorf2.py:349-433 applies no ISCO / f_max truncation (the docstring at :352 says otherwise): a 90 Msun binary
    still has signal in the last bin of a 1024 Hz band, far above its ISCO (~50 Hz)."""
    cfg, like, ref = _pair("C1")
    p = _point(cfg, mchirp=39.0, q=1.0)
    h = like.project_fdwave(dict(p))
    top = np.abs(h["H1"][-1])
    assert np.isfinite(top) and top > 0.0
    assert _close(like.log_like(dict(p)), ref.log_like(dict(p)))


def test_quirk04_tides_only_if_a_lambda_is_nonzero():
    """taylorf2.py:394: the tidal (and quadrupole-monopole, :409-414) phase enters only if lambda1 != 0 or lambda2 != 0;
    the point-mass phase is always evaluated with Lam = dLam = 0 (:385-390)."""
    cfg, like, ref = _pair(approx="TaylorF2_5.5PN_3.5PNQM_7.5PNTides")
    _, like_pm, _ = _pair(approx="TaylorF2_5.5PN")
    zero = _point(cfg, lambda1=0.0, lambda2=0.0)
    tiny = _point(cfg, lambda1=0.0, lambda2=0.5)
    assert like.log_like(dict(zero)) == like_pm.log_like(dict(zero))          # no tidal / QM term at all
    assert _close(like.log_like(dict(zero)), ref.log_like(dict(zero)))
    assert _close(like.log_like(dict(tiny)), ref.log_like(dict(tiny)))


def test_quirk05_complete_tides_use_different_p4_polynomials_for_the_two_bodies():
    """taylorf2.py:115-116: in PhifT7hPNComplete p4a uses (27719, 22415, 7598, 10520) and p4b (27719, 22127, 7022, 10232):
    at q = 1 swapping the two tidal deformabilities changes the 2020 model but not the symmetric 7.5PN one."""
    for approx, symmetric in (("TaylorF2_5.5PN_7.5PNTides", True), ("TaylorF2_5.5PN_7.5PNTides2020", False)):
        cfg, like, ref = _pair(approx=approx)
        a = _point(cfg, q=1.0, lambda1=0.0, lambda2=5000.0)
        b = _point(cfg, q=1.0, lambda1=5000.0, lambda2=0.0)
        la, lb = like.log_like(dict(a)), like.log_like(dict(b))
        assert _close(la, ref.log_like(dict(a))) and _close(lb, ref.log_like(dict(b)))
        tol = harness.tolerance(np.array([la]))[0]
        if symmetric:
            assert abs(la - lb) <= tol
        else:
            assert abs(la - lb) > 5 * tol        # 7 x tol in the oracle: an asymmetry of the MODEL, reproduced


def test_quirk06_55pn_phase_takes_v_without_abs():
    """taylorf2.py:252: Phif5hPN takes (pi M f gt)^(1/3) without abs (every other function uses abs, e.g. :178): a
    negative total mass is NaN -> -inf in the 5.5PN family, finite in the 3.5PN model."""
    cfg, like35, ref35 = _pair("C1")
    _, like55, ref55 = _pair("C1", approx="TaylorF2_5.5PN")
    names = [n if n != "mchirp" else "mtot" for n in NAMES]
    X, _ = syn.draw_walkers(cfg, 4, 3, "narrow")
    X[:, 0] = -60.0
    const = syn.constants(cfg)
    with np.errstate(all="ignore"):
        r35 = harness.eval_batch(ref35, X, names, const)
        r55 = harness.eval_batch(ref55, X, names, const)
    g35, g55 = like35.log_like_batch(X, names, const), like55.log_like_batch(X, names, const)
    assert np.all(np.isfinite(r35)) and np.all(np.isneginf(np.where(np.isnan(r55), -np.inf, r55)))
    assert np.all(np.isneginf(g55)) and np.all(np.abs(g35 - r35) <= harness.tolerance(r35))


def test_quirk07_equal_masses_give_x_exactly_zero():
    """taylorf2.py:30-32, obs/gw/utils/__init__.py:151-154: m1 = (1 + delta)/2, m2 = (1 - delta)/2 with
    delta = sqrt(1 - 4 eta); at q = 1 delta = X = 0 exactly and every tidal combination stays finite."""
    for approx in ("TaylorF2_3.5PN", "TaylorF2_5.5PN_7.5PNTides2020"):
        cfg, like, ref = _pair(approx=approx)
        p = _point(cfg, q=1.0)
        got = like.log_like(dict(p))
        assert np.isfinite(got) and _close(got, ref.log_like(dict(p)))


def test_quirk08_mtot_from_mchirp_and_iota_from_cos_iota():
    """waveform.py:334-336,374-375: mtot = mchirp / nu^0.6 with nu = q / (1 + q)^2, iota = arccos(cos_iota): the same
    point given as (mtot, iota) evaluates to the same logL."""
    cfg, like, ref = _pair("C1")
    p = _point(cfg)
    q = p["q"]
    alt = dict(p)
    alt["mtot"] = alt.pop("mchirp") / (q / (1.0 + q) ** 2.0) ** 0.6
    alt["iota"] = float(np.arccos(alt.pop("cos_iota")))
    a, b = like.log_like(dict(p)), like.log_like(dict(alt))
    assert _close(a, b) and _close(a, ref.log_like(dict(p)))


def test_quirk10_band_mask_is_inclusive_on_both_ends():
    """strain.py:375-379: the in-band axis is freqs[(f >= f_min) & (f <= f_max)]: both end bins belong to it."""
    cfg, like, ref = _pair("C1")
    n = int(round((cfg["f_max"] - cfg["f_min"]) * cfg["seglen"])) + 1
    assert like.engine.n_bins == n == len(ref.dets["H1"].freqs)
    assert ref.dets["H1"].freqs[0] == cfg["f_min"] and ref.dets["H1"].freqs[-1] == cfg["f_max"]


def test_quirk12_roq_weights_use_the_psd_without_the_window_factor():
    """pipe/__init__.py:763 vs detector.py:483: the ROQ weights are built from Noise.interp_psd_pad(f) WITHOUT the
    Series' window factor, the full-grid PSD with it.  (CPU set-up code; window_factor = 1 for the frequency-domain
    Series of the --fd-inject path, strain.py:248, so the product's set-up is checked with a factor forced to 2.)"""
    from conftest import load_golden, roq_product
    gold = load_golden("roq_small.npz")
    cfg, net, pnet, roq = roq_product(gold)
    ifos, datas, dets, noises, f = pnet
    for ifo in ifos:
        datas[ifo].window_factor = 2.0
    import tempfile
    from bajes_b200 import roq as roqmod
    from bajes_b200.ensemble import BoxPrior
    prior = BoxPrior(list(NAMES), [[-1.0, 1.0]] * len(NAMES), {})
    prior.bounds[NAMES.index("time_shift")] = [float(gold["t_bounds"][0]), float(gold["t_bounds"][1])]
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, syn.make_roq_basis(cfg, n_lin=48, n_quad=24, seed=7), cfg)
        roq2 = roqmod.build_roq_dict(tmp, int(gold["time_points"]), ifos, datas, noises, f, cfg["f_min"], cfg["f_max"],
                                     cfg["seglen"], prior, cfg["approx"])
    for ifo in ifos:
        np.testing.assert_array_equal(roq2[ifo]["psi_weights"], roq[ifo]["psi_weights"])
        dets[ifo].store_measurement(datas[ifo], noises[ifo])
        np.testing.assert_allclose(dets[ifo].psd, 2.0 * noises[ifo].interp_psd_pad(dets[ifo].freqs), rtol=1e-15)


def test_quirk13_unphysical_points_return_minus_infinity_not_an_exception():
    """log_like.py:121-129: an all-zero / empty hp, or any NaN / Inf in hp or hc, returns -inf.  The base class' NaN -> RuntimeError (inf/likelihood.py:29-31) is not
    applied by the GW likelihood.  (The survey's example q < 1 is not one: eta = q / (1 + q)^2 <= 1/4 for every real q.)"""
    cfg, like, ref = _pair("C1")
    X, _ = syn.draw_walkers(cfg, 6, 5, "narrow")
    X[0, NAMES.index("q")] = np.nan
    X[1, NAMES.index("distance")] = np.inf
    X[2, NAMES.index("mchirp")] = np.nan
    X[3, NAMES.index("cos_iota")] = 1.5
    X[4, NAMES.index("distance")] = 0.0
    const = syn.constants(cfg)
    got = like.log_like_batch(X, NAMES, const)
    with np.errstate(all="ignore"):
        ref_v = harness.eval_batch(ref, X, NAMES, const)
    assert np.all(np.isneginf(got[:5])) and np.isfinite(got[5])
    assert np.all(np.isneginf(ref_v[:5])) and _close(got[5], ref_v[5])
    p = {k: float(v) for k, v in zip(NAMES, X[0])}
    p.update(const)
    assert like.log_like(dict(p)) == -np.inf


def test_quirk15_inputs_are_read_only_and_derived_keys_are_tolerated():
    """waveform.py:334-377: compute_hphc MUTATES the caller's dict (adds mtot, iota, cartesian spins ...).  Samplers
    pass a fresh dict per call (inf/likelihood.py:55), so the batched path treats its inputs as read-only; a dict that
    already carries the keys the reference's mutation adds evaluates to the same value (the primary parameters win)."""
    cfg, like, ref = _pair("C1")
    p = _point(cfg)
    before = dict(p)
    a = like.log_like(p)
    assert p == before
    mutated = dict(p)
    ref.wave.compute_hphc(mutated)                   # the oracle keeps the mutation of the reference
    assert set(mutated) > set(before)
    assert like.log_like(mutated) == a
