"""
GPU parity tests (run with ``-m gpu`` on a B200).  Every call goes through the C ABI
(``libbajes_b200.so`` via ctypes).  Checked against
  * the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  * the oracle port evaluated live on the host,
with the tolerance of BASELINE.json: |dlogL| <= 1e-6 |logL| + 1e-4 per point.
"""
import numpy as np
import pytest

from bajes_b200 import GWLikelihood, Waveform, synthetic as syn
from oracle import harness

from conftest import load_golden, port_network, product_network_from_port, roq_product

pytestmark = pytest.mark.gpu

TOL_NOTE = "tolerance 1e-6*|logL| + 1e-4 (BASELINE.json north_star)"


def _names(gold):
    return [str(n) for n in gold["names"]]


def _check(got, ref, what, frac=1.0):
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), "%s: finite pattern differs" % what
    assert np.all(np.isneginf(got[~fin])), "%s: unphysical rows must be -inf" % what
    err = np.abs(got[fin] - ref[fin])
    tol = harness.tolerance(ref[fin])
    worst = float(np.max(err / tol))
    print("%s: max |dlogL|/tol = %.3g over %d points (logL in [%.1f, %.1f])"
          % (what, worst, fin.sum(), ref[fin].min(), ref[fin].max()))
    assert worst <= frac, "%s: %g x tol; %s" % (what, worst, TOL_NOTE)
    return worst


def _like(cfg, gold, net=None, **kw):
    net = net if net is not None else port_network(cfg, gold)
    ifos, datas, dets, noises, f = product_network_from_port(cfg, net)
    approx = kw.pop("approx", cfg["approx"])
    return GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], approx, **kw)


# ------------------------------------------------------------------------------------------------------
# config 1: TaylorF2 3.5PN BBH, H1+L1, 4 s, 20-1024 Hz
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sincos", ["mufu", "poly32", "fp64"])
@pytest.mark.parametrize("marg", [False, True])
def test_c1_vs_reference_golden(sincos, marg):
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    like = _like(cfg, gold, marg_phi_ref=marg, sincos=sincos)
    got = like.log_like_batch(gold["X"], _names(gold), syn.constants(cfg))
    _check(got, gold["logl_marg%d" % int(marg)], "C1 %s marg=%s" % (sincos, marg))


def test_c1_per_detector_inner_products_and_single_point_api():
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    like = _like(cfg, gold, sincos="fp64")
    names, const = _names(gold), syn.constants(cfg)
    batch = like.log_like_batch(gold["X"], names, const)
    for r in range(gold["parts"].shape[0]):
        p = {k: float(v) for k, v in zip(names, gold["X"][r])}
        p.update(const)
        # dict-in / float-out contract of GWLikelihood.log_like, bit-identical to the batched entry
        assert like.log_like(dict(p)) == batch[r]
        ip = like.inner_products(p)
        for i, ifo in enumerate(cfg["ifos"]):
            dh, hh = ip[ifo]
            ref = gold["parts"][r, i]
            scale = np.hypot(ref[0], ref[1]) + np.sqrt(ref[2])
            assert abs(dh.real - ref[0]) <= 1e-6 * scale + 1e-5
            assert abs(dh.imag - ref[1]) <= 1e-6 * scale + 1e-5
            assert abs(hh - ref[2]) <= 1e-10 * ref[2]        # <h|h> is phase-free, pure FP64
    # unphysical rows: -inf, never an exception (log_like.py:121-129)
    bad = {k: float(v) for k, v in zip(names, gold["X"][-1])}
    bad.update(const)
    assert like.log_like(bad) == -np.inf


def test_c1_projected_waveform_matches_oracle():
    """Waveform.compute_hphc + Detector.project_fdwave per bin (FP64 on the GPU) vs the oracle port."""
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    net = port_network(cfg, gold)
    like = _like(cfg, gold, net=net)
    port = harness.port_likelihood(cfg, net=net)
    names, const = _names(gold), syn.constants(cfg)
    for r in (0, 5, 70):
        p = {k: float(v) for k, v in zip(names, gold["X"][r])}
        p.update(const)
        ours = like.project_fdwave(dict(p))
        wave = port.wave.compute_hphc(dict(p))
        for ifo in cfg["ifos"]:
            ref = port.dets[ifo].project_fdwave(wave, p)
            # phases reach ~1e4 rad: agreement is limited by rounding of the phase, ~1e-11 relative
            assert np.max(np.abs(ours[ifo] - ref)) <= 1e-9 * np.max(np.abs(ref))
    # the drop-in Waveform object
    band = gold["band"]
    w = Waveform(gold["freqs"][band], cfg["srate"], cfg["seglen"], cfg["approx"])
    p = syn.injection_params(cfg)
    hphc = w.compute_hphc(dict(p))
    ref = port.wave.compute_hphc(dict(p))
    assert np.max(np.abs(hphc.plus - ref.plus)) <= 1e-9 * np.max(np.abs(ref.plus))
    assert np.max(np.abs(hphc.cross - ref.cross)) <= 1e-9 * np.max(np.abs(ref.cross))


# ------------------------------------------------------------------------------------------------------
# reduced config 2 (16 s, 3 detectors): every approximant of the TaylorF2 family
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("approx,marg,n", [
    ("TaylorF2_3.5PN", False, 256), ("TaylorF2_3.5PN", True, 64), ("TaylorF2_5.5PN", False, 64),
    ("TaylorF2_5.5PN_7.5PNTides", False, 64), ("TaylorF2_5.5PN_3.5PNQM_7.5PNTides", False, 64),
    ("TaylorF2_5.5PN_7.5PNTides2020", False, 64)])
def test_c2_small_all_approximants(approx, marg, n):
    cfg = syn.CONFIGS["C2_small"]
    gold = load_golden("c2_small.npz")
    like = _like(cfg, gold, approx=approx, marg_phi_ref=marg)
    got = like.log_like_batch(gold["X"][:n], _names(gold), syn.constants(cfg))
    _check(got, gold["logl_%s_marg%d" % (approx, int(marg))][:n], "C2_small %s marg=%s" % (approx, marg))


# ------------------------------------------------------------------------------------------------------
# config 2 at full size: 128 s, 521 729 bins, H1/L1/V1
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_full():
    cfg = syn.CONFIGS["C2"]
    gold = load_golden("c2.npz")
    net = port_network(cfg, gold)
    return cfg, gold, net


@pytest.mark.parametrize("sincos", ["mufu", "poly32"])
def test_c2_full_vs_reference_golden(c2_full, sincos):
    cfg, gold, net = c2_full
    like = _like(cfg, gold, net=net, sincos=sincos)
    got = like.log_like_batch(gold["X"], _names(gold), syn.constants(cfg))
    _check(got, gold["logl_TaylorF2_3.5PN_marg0"], "C2 full %s" % sincos)


def test_c2_full_55pn_tides(c2_full):
    cfg, gold, net = c2_full
    like = _like(cfg, gold, net=net, approx="TaylorF2_5.5PN_7.5PNTides")
    ref = gold["logl_TaylorF2_5.5PN_7.5PNTides_marg0"]
    got = like.log_like_batch(gold["X"][:len(ref)], _names(gold), syn.constants(cfg))
    _check(got, ref, "C2 full 5.5PN+7.5PN tides")


def test_c2_full_size_properties(c2_full):
    """Size-independent properties at the full 4096-walker batch (the oracle would need ~15 min here)."""
    cfg, gold, net = c2_full
    like = _like(cfg, gold, net=net)
    names, const = _names(gold), syn.constants(cfg)
    X, _ = syn.draw_walkers(cfg, 4096, 77, "narrow")
    full = like.log_like_batch(X, names, const)
    assert np.all(np.isfinite(full))
    # (a) determinism: run-to-run identical bits
    assert np.array_equal(full, like.log_like_batch(X, names, const))
    # (b) sharding invariance: any split of the walker batch gives the same bits per walker
    parts = np.concatenate([like.log_like_batch(X[:1000], names, const), like.log_like_batch(X[1000:1003], names, const),
                            like.log_like_batch(X[1003:], names, const)])
    assert np.array_equal(full, parts)
    # (c) distance scaling: <d|h> ~ 1/D, <h|h> ~ 1/D^2  =>  logL(D) = a/D - b/D^2
    i_d = names.index("distance")
    X2, X3 = X[:64].copy(), X[:64].copy()
    X2[:, i_d] *= 2.0
    X3[:, i_d] *= 4.0
    l1, l2, l3 = full[:64], like.log_like_batch(X2, names, const), like.log_like_batch(X3, names, const)
    # l1 = a - b, l2 = a/2 - b/4, l3 = a/4 - b/16  =>  l1 - 6 l2 + 8 l3 = 0
    resid = l1 - 6.0 * l2 + 8.0 * l3
    assert np.all(np.abs(resid) <= 3 * harness.tolerance(l1) + 1e-9 * np.abs(l1)), np.max(np.abs(resid))
    # (d) phi_ref shift by pi flips the sign of <d|h>: logL(phi) + logL(phi + pi) = -<h|h>
    i_p = names.index("phi_ref")
    Xp = X[:64].copy()
    Xp[:, i_p] += np.pi
    lp = like.log_like_batch(Xp, names, const)
    for r in range(4):
        p = {k: float(v) for k, v in zip(names, X[r])}
        p.update(const)
        hh = sum(v[1] for v in like.inner_products(p).values())
        assert abs((l1[r] + lp[r]) + hh) <= 2 * harness.tolerance(l1[r]) + 1e-6 * hh


@pytest.mark.parametrize("approx", ["TaylorF2_3.5PN", "TaylorF2_5.5PN_7.5PNTides"])
def test_c2_full_size_production_kernel_vs_fp64_kernel(c2_full, approx):
    """8192 walkers (posterior-scale and prior-scale boxes) x 521 729 bins: the tensor-core mixed-precision kernel
    against the all-FP64 validation kernel, itself pinned to the reference on the golden subset.  Catches rare
    parameter-dependent failures the 256 golden points could miss."""
    cfg, gold, net = c2_full
    names, const = _names(gold), syn.constants(cfg)
    Xa, _ = syn.draw_walkers(cfg, 4096, 91, "narrow")
    Xb, _ = syn.draw_walkers(cfg, 4096, 92, "wide")
    X = np.concatenate([Xa, Xb])
    fast = _like(cfg, gold, net=net, approx=approx).log_like_batch(X, names, const)
    slow = _like(cfg, gold, net=net, approx=approx, sincos="fp64").log_like_batch(X, names, const)
    _check(fast, slow, "C2 full, 8192 walkers, %s: production vs FP64 kernel" % approx)


@pytest.mark.parametrize("sincos", ["mufu", "poly32", "fp64"])
def test_wide_phase_path_subsolar_mass(sincos):
    """Sub-solar-mass binaries accumulate > 2^19 turns of phase, beyond the range of the binary-angle trick: the
    prologue flags them and their walker tile takes the exact-reduction path.  Mixed into an ordinary batch."""
    cfg = syn.CONFIGS["C2_small"]
    gold = load_golden("c2_small.npz")
    net = port_network(cfg, gold)
    like = _like(cfg, gold, net=net, sincos=sincos)
    port = harness.port_likelihood(cfg, net=net)
    names, const = _names(gold), syn.constants(cfg)
    X = gold["X"][:40].copy()
    rng = np.random.default_rng(5)
    rows = [1, 7, 8, 30]
    X[rows, names.index("mchirp")] = rng.uniform(0.02, 0.035, len(rows))
    X[rows, names.index("distance")] = 1.0
    got = like.log_like_batch(X, names, const)
    want = harness.eval_batch(port, X, names, const)
    _check(got, want, "wide-phase path %s" % sincos)
    # the ordinary walkers of the batch are untouched by their neighbours taking the slow path ...
    others = [r for r in range(40) if r not in rows and r // 128 != rows[0] // 128]
    # ... (all 40 share one 128-walker tile here, so compare against the golden values instead)
    ref = gold["logl_TaylorF2_3.5PN_marg0"][:40]
    keep = np.array([r not in rows for r in range(40)])
    assert np.all(np.abs(got[keep] - ref[keep]) <= harness.tolerance(ref[keep]))


def test_zero_noise_injection_peak():
    """Noise-free injection: logL at the true parameters equals <h|h>/2."""
    cfg = dict(syn.CONFIGS["C2_small"])
    cl = syn.product_classes()
    ifos, datas, dets, noises, f = syn.build_network(cfg, cl, inject=True, noise=False)
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64")
    p = syn.injection_params(cfg)
    hh = sum(v[1] for v in like.inner_products(dict(p)).values())
    assert abs(like.log_like(dict(p)) - 0.5 * hh) <= 1e-6 * hh


# ------------------------------------------------------------------------------------------------------
# ROQ (config 3 at reduced size): linear + quadratic weights in JenpyROQ format
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("marg", [False, True])
def test_roq_vs_reference_golden(marg):
    # the roq dict comes from the PRODUCT set-up (bajes_b200.roq.build_roq_dict), not from the oracle's
    gold = load_golden("roq_small.npz")
    cfg, net, (ifos, datas, dets, noises, f), roq = roq_product(gold)
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"],
                        marg_phi_ref=marg, roq=roq)
    got = like.log_like_batch(gold["X"], _names(gold), syn.constants(cfg))
    ref = gold["logl_marg%d" % int(marg)]
    ref = np.where(np.isfinite(ref), ref, -np.inf)     # reference yields nan/-inf outside the tc grid
    _check(got, ref, "ROQ marg=%s" % marg)


# ------------------------------------------------------------------------------------------------------
# relative binning (config 4 at reduced size).  PARITY UNPINNED against the reference (dead code upstream):
# checked against the NumPy restatement, and for accuracy against the full-grid likelihood.
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("marg", [False, True])
def test_relbin_vs_port_restatement(marg):
    from bajes_b200.binning import GWBinningLikelihood
    from oracle import bajes_port as port
    cfg = syn.CONFIGS["C4_small"]
    cl = harness.port_classes()
    net = syn.build_network(cfg, cl)
    fid = syn.injection_params(cfg)
    ref = port.GWBinningLikelihoodPort(net[0], net[1], net[2], net[3], dict(fid), net[4], cfg["srate"], cfg["seglen"],
                                       cfg["approx"], marg_phi_ref=marg)
    # fresh detector objects for the product (the port re-points its detectors at the bin edges)
    net2 = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = product_network_from_port(cfg, net2)
    like = GWBinningLikelihood(ifos, datas, dets, noises, dict(fid), f, cfg["srate"], cfg["seglen"], cfg["approx"],
                               marg_phi_ref=marg)
    assert like.Nbin == ref.Nbin and np.array_equal(like.fbin_ind, ref.fbin_ind)
    np.testing.assert_allclose(like.sdat, ref.sdat, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref.sdat)))
    xa, names = syn.draw_walkers(cfg, 1500, 41, "narrow")
    xb, _ = syn.draw_walkers(cfg, 500, 42, "wide")
    X = np.concatenate([xa, xb])
    const = syn.constants(cfg)
    got = like.log_like_batch(X, names, const)
    want = harness.eval_batch(ref, X, names, const)
    _check(got, want, "relbin, 2000 points, marg=%s" % marg)
    p = {k: float(v) for k, v in zip(names, X[3])}
    p.update(const)
    assert like.log_like(dict(p)) == got[3]


def test_relbin_tracks_full_grid_likelihood():
    """The method's own accuracy: near the fiducial point relative binning reproduces the full-grid logL."""
    from bajes_b200.binning import GWBinningLikelihood
    cfg = syn.CONFIGS["C4_small"]
    cl = syn.product_classes()
    fid = syn.injection_params(cfg)
    full = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64")
    rb = GWBinningLikelihood(*syn.build_network(cfg, cl)[:4], dict(fid), syn.frequency_axis(cfg), cfg["srate"],
                             cfg["seglen"], cfg["approx"])
    X, names = syn.draw_walkers(cfg, 64, 43, "narrow")
    const = syn.constants(cfg)
    a = full.log_like_batch(X, names, const)
    b = rb.log_like_batch(X, names, const)
    at_fid = abs(rb.log_like(dict(fid)) - full.log_like(dict(fid)))
    print("relbin vs full grid: |dlogL| at fiducial %.2e, max over narrow box %.3g (logL drop up to %.0f)"
          % (at_fid, np.max(np.abs(a - b)), np.max(a) - np.min(a)))
    assert at_fid < 5e-2           # the bin slices exclude the last in-band sample (binning.py:67-69)
    assert np.max(np.abs(a - b)) <= 0.02 * (np.max(a) - np.min(a)) + 0.05


# ------------------------------------------------------------------------------------------------------
# calibration envelopes + PSD weights (detector.py:517-531), against reference golden vectors
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sincos", ["mufu", "poly32", "fp64"])
@pytest.mark.parametrize("key", ["spcal_marg0", "weights_marg0", "both_marg0", "both_marg1"])
def test_calibration_envelopes_and_psd_weights(key, sincos):
    cfg = syn.CONFIGS["C2_small"]
    gold = load_golden("extras_small.npz")
    kw = {}
    if "spcal_freqs_" + key in gold:
        kw.update(nspcal=len(gold["spcal_freqs_" + key]), spcal_freqs=gold["spcal_freqs_" + key])
    if "len_weights_" + key in gold:
        kw.update(nweights=len(gold["len_weights_" + key]), len_weights=gold["len_weights_" + key])
    like = _like(cfg, gold, marg_phi_ref=key.endswith("1"), sincos=sincos, **kw)
    names = [str(n) for n in gold["names_" + key]]
    X, const = gold["X_" + key], syn.constants(cfg)
    got = like.log_like_batch(X, names, const)
    _check(got, gold["logl_" + key], "extras %s %s" % (key, sincos))
    # dict-in / float-out contract with the extra parameters in the dict
    p = {k: float(v) for k, v in zip(names, X[2])}
    p.update(const)
    assert like.log_like(dict(p)) == got[2]
    # parameters may also arrive as constants (e.g. a fixed calibration node)
    n0 = len(syn.WALKER_NAMES)
    const2 = dict(const)
    const2.update({k: float(v) for k, v in zip(names[n0:], X[5, n0:])})
    one = like.log_like_batch(X[5:6, :n0], names[:n0], const2)
    assert one[0] == got[5]


def test_detector_inner_products_with_psd_weights_are_consistent():
    """Detector.compute_inner_products with PSD weights (detector.py:524-531): dd is recomputed with psd * weights and
    the fourth value is sum(log w); together with the weighted dh / hh they rebuild logL as log_like.py:181 does."""
    cfg = syn.CONFIGS["C2_small"]
    gold = load_golden("extras_small.npz")
    key = "both_marg0"
    like = _like(cfg, gold, nspcal=len(gold["spcal_freqs_" + key]), spcal_freqs=gold["spcal_freqs_" + key],
                 nweights=len(gold["len_weights_" + key]), len_weights=gold["len_weights_" + key])
    names = [str(n) for n in gold["names_" + key]]
    p = {k: float(v) for k, v in zip(names, gold["X_" + key][3])}
    p.update(syn.constants(cfg))
    dh = hh = dd = w = 0.0
    for ifo in like.ifos:
        a, b, c, d = like.dets[ifo].compute_inner_products(None, dict(p), "freq", psd_weight_factor=True)
        assert d != 0.0 and c != like.dets[ifo]._dd
        dh, hh, dd, w = dh + a, hh + b, dd + c, w + d
    rebuilt = -0.5 * (hh + dd) + np.real(dh) - like.logZ_noise - 0.5 * w
    ref = gold["logl_" + key][3]
    assert abs(rebuilt - ref) <= harness.tolerance(np.array([ref]))[0]


def test_time_marginalised_context_parts_and_frequency_split():
    """A time-marginalised likelihood still reports the per-detector <d|h>, <h|h> of a point (they do not depend on
    the marginalisation), and refuses the frequency split (the FFT needs a walker's whole spectrum on one GPU)."""
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    plain = _like(cfg, gold)
    tm = _like(cfg, gold, marg_time_shift=True)
    p = {k: float(v) for k, v in zip(_names(gold), gold["X"][1])}
    p.update(syn.constants(cfg))
    a, b = plain.inner_products(dict(p)), tm.inner_products(dict(p))
    for ifo in plain.ifos:          # same numbers up to FP32 rounding (the plain context evaluates its heaviest bins in FP64)
        assert abs(a[ifo][0] - b[ifo][0]) <= 1e-6 * abs(a[ifo][0]) and abs(a[ifo][1] - b[ifo][1]) <= 1e-9 * a[ifo][1]
    assert np.isfinite(tm.log_like(dict(p))) and tm.log_like(dict(p)) != plain.log_like(dict(p))
    import torch
    x = torch.zeros((1, 13), dtype=torch.float64, device="cuda")
    sums = torch.zeros((4, 1), dtype=torch.float64, device="cuda")
    with pytest.raises(NotImplementedError):
        tm.engine.partial_device(x.data_ptr(), 1, 13, tm.param_map(_names(gold), syn.constants(cfg)), 0, 1,
                                 sums.data_ptr())


# ------------------------------------------------------------------------------------------------------
# time-shift marginalisation (log_like.py:135-156), against reference golden vectors
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sincos", ["mufu", "poly32"])
@pytest.mark.parametrize("name", ["C1", "C2_small"])
@pytest.mark.parametrize("mphi", [False, True])
def test_time_marginalised_likelihood(name, mphi, sincos):
    cfg = syn.CONFIGS[name]
    gold = load_golden("tmarg.npz")
    cl = harness.port_classes(gmst_reference=float(gold["gmst_" + name][0]), sday=float(gold["sday_" + name][0]))
    net = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = product_network_from_port(cfg, net)
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"],
                        marg_time_shift=True, marg_phi_ref=mphi, sincos=sincos)
    names = [str(n) for n in gold["names"]]
    X, const = gold["X_" + name], syn.constants(cfg)
    got = like.log_like_batch(X, names, const)
    _check(got, gold["logl_%s_mphi%d" % (name, int(mphi))], "tmarg %s mphi=%s %s" % (name, mphi, sincos))
    p = {k: float(v) for k, v in zip(names, X[1])}
    p.update(const)
    assert like.log_like(dict(p)) == got[1]


def test_engine_reports_launch_and_timing():
    cfg = syn.CONFIGS["C1"]
    gold = load_golden("c1_bbh.npz")
    like = _like(cfg, gold)
    like.log_like_batch(gold["X"], _names(gold), syn.constants(cfg))
    ms = like.engine.last_kernel_ms()
    info = like.engine.last_launch_info()
    assert all(m >= 0 for m in ms) and ms[1] > 0
    # info: grid.x, grid.y, threads, smem of the fast launch; bins per chunk, chunks in all (fast + FP64 window), ...
    assert info[2] in (128, 256) and info[0] <= info[5] and info[5] * info[4] >= int(gold["band"].sum())


# ------------------------------------------------------------------------------------------------------
# the three production code paths of the full grid: tensor-core kernel with integer ramps (default), tensor-core kernel
# with FP64 ramps (non-uniform grids), Horner kernel (calibration envelopes / BAJES_B200_NO_TC)
# ------------------------------------------------------------------------------------------------------
VARIANT_ENVS = ("BAJES_B200_NO_IRAMP", "BAJES_B200_NO_TC", "BAJES_B200_NO_ROT", "BAJES_B200_PRECISE_FRAC", "BAJES_B200_PRECISE_SHARE",
                "BAJES_B200_WINDOW_GENERIC")


@pytest.mark.parametrize("env", [{}, {"BAJES_B200_NO_IRAMP": "1"}, {"BAJES_B200_NO_TC": "1"}, {"BAJES_B200_NO_ROT": "1"},
                                 {"BAJES_B200_PRECISE_FRAC": "0"}, {"BAJES_B200_PRECISE_SHARE": "0.99", "BAJES_B200_PRECISE_FRAC": "0.5"},
                                 {"BAJES_B200_WINDOW_GENERIC": "1"}])
@pytest.mark.parametrize("approx", ["TaylorF2_3.5PN", "TaylorF2_5.5PN_7.5PNTides"])
def test_kernel_variants_vs_reference_golden(env, approx, monkeypatch):
    for k in VARIANT_ENVS:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)                     # read when the context is created
    cfg = syn.CONFIGS["C2_small"]
    gold = load_golden("c2_small.npz")
    like = _like(cfg, gold, approx=approx)
    n = 64
    got = like.log_like_batch(gold["X"][:n], _names(gold), syn.constants(cfg))
    _check(got, gold["logl_%s_marg0" % approx][:n], "C2_small %s %s" % (approx, env or "default"))
    threads = like.engine.last_launch_info()[2]
    assert threads == (128 if "BAJES_B200_NO_TC" in env else 256)


ALL_APPROX = ["TaylorF2_3.5PN", "TaylorF2_5.5PN", "TaylorF2_5.5PN_7.5PNTides", "TaylorF2_5.5PN_3.5PNQM_7.5PNTides",
              "TaylorF2_5.5PN_7.5PNTides2020"]


@pytest.mark.parametrize("name,approx,sincos",
                         [("C2_small", a, "mufu") for a in ALL_APPROX] + [("C2_small", "TaylorF2_3.5PN", "poly32"),
                                                                         ("C1", "TaylorF2_3.5PN", "mufu"),
                                                                         ("C1", "TaylorF2_5.5PN", "poly32")])
def test_hundred_thousand_walkers_vs_fp64_kernel(name, approx, sincos):
    """100 003 walkers (posterior-scale + prior-scale boxes) against the all-FP64 validation kernel, every approximant
    in the default mode, gated at 0.6 of the tolerance (scripts/stress_tail.py runs the same comparison with 10^7
    walkers per approximant: profiles/r02_stress_tail_1e7.txt).  At this count the rare parameter combinations show up
    -- e.g. heavy binaries whose 3.5PN amplitude series changes sign inside the band (bajes applies no ISCO cut): the
    per-stage amplitude correction used to blow up there (500 x the tolerance for 1 walker in 20 000) -- and so does
    the worst case of the FP32 rounding noise, which the FP64 window over the heaviest bins keeps in check."""
    cfg = syn.CONFIGS[name]
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    W = 100003
    Xa, names = syn.draw_walkers(cfg, W // 2, 11, "narrow")
    Xb, _ = syn.draw_walkers(cfg, W - W // 2, 12, "wide")
    X = np.concatenate([Xa, Xb])
    const = syn.constants(cfg)
    slow = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64").log_like_batch(X, names, const)
    fast = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos=sincos).log_like_batch(X, names, const)
    _check(fast, slow, "%s %s, 100 003 walkers, %s vs FP64 kernel" % (name, approx, sincos), frac=0.6)


# ------------------------------------------------------------------------------------------------------
# many points against the oracle port evaluated live on the host (the golden files hold a few dozen points each)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("marg", [False, True])
def test_roq_many_points_vs_port(marg):
    from test_oracle_golden import _roq_port
    gold = load_golden("roq_small.npz")
    cfg, net, roq = _roq_port(gold)
    port_like = harness.port_likelihood(cfg, marg_phi_ref=marg, roq=roq, net=net)
    # product side: its own set-up, with the weight matrices evaluated on the GPU
    _, _, (ifos, datas, dets, noises, f), roq_p = roq_product(gold, device="cuda:0")
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"],
                        marg_phi_ref=marg, roq=roq_p)
    Xa, names = syn.draw_walkers(cfg, 1500, 31, "narrow")
    Xb, _ = syn.draw_walkers(cfg, 1500, 32, "wide")
    X = np.concatenate([Xa, Xb])
    # keep the time shifts inside the tabulated tc range most of the time, leave some outside (-> -inf)
    it = names.index("time_shift")
    t_lo, t_hi = gold["t_bounds"]
    X[:, it] = np.random.default_rng(33).uniform(t_lo - 0.06, t_hi + 0.06, len(X))
    const = syn.constants(cfg)
    with np.errstate(all="ignore"):
        ref = harness.eval_batch(port_like, X, names, const)
    ref = np.where(np.isfinite(ref), ref, -np.inf)
    got = like.log_like_batch(X, names, const)
    assert np.isfinite(ref).sum() > 1000 and np.isneginf(ref).sum() > 10
    _check(got, ref, "ROQ, 3000 points, marg=%s" % marg)


@pytest.mark.parametrize("mphi", [False, True])
def test_time_marginalised_many_points_vs_port(mphi):
    from oracle import bajes_port as port
    cfg = syn.CONFIGS["C1"]
    cl = harness.port_classes()
    net = syn.build_network(cfg, cl)
    ref_like = port.GWLikelihoodPort(net[0], net[1], net[2], net[3], net[4], cfg["srate"], cfg["seglen"], cfg["approx"],
                                     marg_time_shift=True, marg_phi_ref=mphi)
    ifos, datas, dets, noises, f = product_network_from_port(cfg, syn.build_network(cfg, cl))
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"],
                        marg_time_shift=True, marg_phi_ref=mphi)
    Xa, names = syn.draw_walkers(cfg, 600, 41, "narrow")
    Xb, _ = syn.draw_walkers(cfg, 600, 42, "wide")
    X = np.concatenate([Xa, Xb])
    const = syn.constants(cfg)
    with np.errstate(all="ignore"):
        ref = harness.eval_batch(ref_like, X, names, const)
    got = like.log_like_batch(X, names, const)
    _check(got, ref, "time-marginalised C1, 1200 points, mphi=%s" % mphi)


@pytest.mark.parametrize("kind,upd", [
    ("C2_small", dict(seglen=4.0)), ("C2_small", dict(f_min=10.0)), ("C2_small", dict(ifos=["H1"])),
    ("C2_small", dict(f_min=23.3, f_max=1777.7)), ("C2_small", dict(seglen=256.0, f_max=1024.0, srate=2048.0)),
    ("C1", dict(seglen=2.0, f_max=512.0, srate=1024.0)), ("C1", dict(seglen=8.0, f_min=10.0))])
def test_data_configurations_vs_fp64_kernel(kind, upd):
    """Segment lengths from 2 to 256 s, bands from 10 Hz, off-grid band edges, one to three detectors: 20 011 walkers
    each, production kernel against the all-FP64 validation kernel."""
    cfg = dict(syn.CONFIGS[kind])
    cfg.update(upd)
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    W = 20011
    Xa, names = syn.draw_walkers(cfg, W // 2, 51, "narrow")
    Xb, _ = syn.draw_walkers(cfg, W - W // 2, 52, "wide")
    X = np.concatenate([Xa, Xb])
    const = syn.constants(cfg)
    slow = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64").log_like_batch(X, names, const)
    fast = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"]).log_like_batch(X, names, const)
    _check(fast, slow, "%s %s, 20 011 walkers vs FP64 kernel" % (kind, upd))


# ------------------------------------------------------------------------------------------------------
# C5 / N1: the call trace of the UNMODIFIED reference ptmcmc sampler (recorded in the build container with the oracle
# behind the adapters, scripts/make_ptmcmc_trace.py) replayed on the GPU likelihood through the same adapters
# ------------------------------------------------------------------------------------------------------
def test_reference_ptmcmc_call_trace_on_the_gpu():
    """Every batch bajes' parallel-tempered sampler sent through pool.map(posterior.log_likeprior, points)
    (ptmcmc.py:272-281; 2 temperatures x 16 walkers x 12 iterations on the 16 s BNS injection) is sent through
    BatchedPool -> Posterior.log_likeprior_batch -> GWLikelihood.log_like_batch on the GPU: same number of calls, same
    batch shapes, and every log-likelihood the sampler saw reproduced within the tolerance."""
    from bajes_b200.likelihood import Posterior
    from bajes_b200.pool import BatchedPool
    cfg = syn.CONFIGS["C2_small"]
    tr = load_golden("ptmcmc_trace.npz")
    like = _like(cfg, tr)
    names = [str(n) for n in tr["names"]]
    const = {str(k): float(v) for k, v in zip(tr["const_keys"], tr["const_vals"])}

    class _OpenPrior(object):               # the sampler already applied the bounds: every recorded point is in the box
        pass

    prior = _OpenPrior()
    prior.names, prior.const = names, const
    prior.in_bounds = lambda x: True
    prior.log_prior = lambda x: 0.0
    post = Posterior(like, prior)
    pool = BatchedPool(post)
    k, worst = 0, 0.0
    for n in tr["sizes"]:
        pts = tr["X"][k:k + n]
        res = pool.map(post.log_likeprior, pts)
        got = np.array([r[0] for r in res])
        worst = max(worst, _check(got, tr["logl"][k:k + n], "ptmcmc call of %d points" % n))
        k += n
    assert pool.n_batched_calls == len(tr["sizes"]) == 1 + 2 * int(tr["shape"][2]) and pool.n_points == len(tr["X"])
    print("reference ptmcmc trace: %d calls, %d points, worst %.3g x tol" % (len(tr["sizes"]), k, worst))
