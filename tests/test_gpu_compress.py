"""
GPU tests of the compressed frequency axis, through the C ABI.  The axis has no counterpart in the
reference; the bar is the reference's: |dlogL| <= 1e-6 |logL| + 1e-4 against its golden vectors / the all-FP64 full-grid
kernel, plus what only this path can get wrong -- routing of walkers outside the design, and bit-reproducibility.
"""
import numpy as np
import pytest

from bajes_b200 import GWLikelihood, synthetic as syn
from oracle import harness

from conftest import load_golden, port_network, product_network_from_port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    cfg = syn.CONFIGS["C2"]
    gold = load_golden("c2.npz")
    net = product_network_from_port(cfg, port_network(cfg, gold))
    names = [str(n) for n in gold["names"]]
    return cfg, gold, net, names, syn.constants(cfg)


def _like(cfg, net, **kw):
    ifos, datas, dets, noises, f = net
    return GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], kw.pop("approx", cfg["approx"]), **kw)


def _ratio(got, ref):
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    return float(np.max(np.abs(got[fin] - ref[fin]) / harness.tolerance(ref[fin])))


def test_context_has_two_levels_and_the_golden_points_pass_on_the_node_axis(c2):
    cfg, gold, net, names, const = c2
    like = _like(cfg, net)
    i0, i1 = like.engine.compress_info(0), like.engine.compress_info(1)
    assert 0 < i0["nodes"] < i0["n_bins"] // 15 and i0["nodes"] < i1["nodes"] < i0["n_bins"] // 5
    assert i0["tau_max"] == pytest.approx(0.11) and i1["tau_max"] == pytest.approx(1.1)
    got = like.log_like_batch(gold["X"], names, const)
    counts = like.engine.last_level_counts()
    assert sum(counts) == len(gold["X"]) and counts[0] > 0
    assert _ratio(got, gold["logl_TaylorF2_3.5PN_marg0"]) <= 1.0
    # and the option that switches it off evaluates every bin with the kernels of round 1
    off = _like(cfg, net, compress=False)
    assert off.engine.compress_info(0)["nodes"] == 0
    assert _ratio(off.log_like_batch(gold["X"], names, const), gold["logl_TaylorF2_3.5PN_marg0"]) <= 1.0
    assert off.engine.last_level_counts() == (0, 0, 0)


def test_walkers_outside_the_design_are_routed_to_the_table_that_covers_them(c2):
    cfg, gold, net, names, const = c2
    comp, full, ref = _like(cfg, net), _like(cfg, net, compress=False), _like(cfg, net, sincos="fp64")
    Xn, _ = syn.draw_walkers(cfg, 300, 41, "narrow")
    Xw, _ = syn.draw_walkers(cfg, 300, 42, "wide")
    X = np.concatenate([Xn, Xw])
    it, im = names.index("time_shift"), names.index("mchirp")
    X[0::6, it] = 0.4            # beyond level 0 (0.11 s), inside level 1 (1.1 s)
    X[1::6, it] = -0.9
    X[2::6, it] = 1.7            # beyond both: full grid
    X[3::6, im] = 0.3            # lighter than both levels (0.8 / 0.4): full grid
    X[4, im] = 0.02              # sub-solar: wide-phase path on the full grid
    got = comp.log_like_batch(X, names, const)
    c0, c1, cf = comp.engine.last_level_counts()
    assert c0 + c1 + cf == len(X)
    assert c1 >= 190 and cf >= 195 and c0 >= 150, (c0, c1, cf)
    assert _ratio(got, ref.log_like_batch(X, names, const)) <= 1.0
    # the walkers that went to the full grid saw every bin (launch_multi evaluates all tables with the non-uniform-axis
    # instantiation of the tile kernel, so their bits differ from a context without compression -- within the tolerance)
    rows = np.r_[np.arange(2, len(X), 6), np.arange(3, len(X), 6), 4]
    g_full = full.log_like_batch(X, names, const)
    assert _ratio(got[rows], g_full[rows]) <= 0.7 and _ratio(got, g_full) <= 1.0
    # any order, any split: same bits (the work lists are filled with atomics)
    perm = np.random.default_rng(0).permutation(len(X))
    again = comp.log_like_batch(X[perm], names, const)
    assert np.array_equal(again, got[perm])
    parts = np.concatenate([comp.log_like_batch(X[:77], names, const), comp.log_like_batch(X[77:], names, const)])
    assert np.array_equal(parts, got)


@pytest.mark.parametrize("approx", ["TaylorF2_3.5PN", "TaylorF2_5.5PN_7.5PNTides2020"])
def test_node_axis_against_the_fp64_full_grid_kernel_on_a_large_batch(c2, approx):
    """16 384 walkers (posterior- and prior-scale boxes): compressed axis vs the all-FP64 kernel on every bin"""
    cfg, gold, net, names, const = c2
    comp, ref = _like(cfg, net, approx=approx), _like(cfg, net, approx=approx, sincos="fp64")
    Xn, _ = syn.draw_walkers(cfg, 8192, 51, "narrow")
    Xw, _ = syn.draw_walkers(cfg, 8192, 52, "wide")
    X = np.concatenate([Xn, Xw])
    r = _ratio(comp.log_like_batch(X, names, const), ref.log_like_batch(X, names, const))
    print("compressed axis vs fp64 full grid, %s, 16384 walkers: %.3f tol; tables %s" % (approx, r, comp.engine.last_level_counts()))
    assert r <= 0.7


def test_inner_products_and_single_point_api_on_the_node_axis(c2):
    cfg, gold, net, names, const = c2
    comp, ref = _like(cfg, net), _like(cfg, net, sincos="fp64")
    x = gold["X"][3]
    p = {k: float(v) for k, v in zip(names, x)}
    p.update(const)
    a, b = comp.inner_products(dict(p)), ref.inner_products(dict(p))
    for ifo in comp.ifos:
        dh_a, hh_a = a[ifo][0], a[ifo][1]
        dh_b, hh_b = b[ifo][0], b[ifo][1]
        assert abs(dh_a - dh_b) <= 1e-4 + 1e-6 * abs(dh_b)
        assert abs(hh_a - hh_b) <= 1e-9 * abs(hh_b)
    assert comp.log_like(dict(p)) == comp.log_like_batch(x.reshape(1, -1), names, const)[0]


@pytest.mark.parametrize("n_ifo,marg,approx", [(1, False, "TaylorF2_3.5PN"), (2, True, "TaylorF2_5.5PN_3.5PNQM_7.5PNTides"),
                                               (3, True, "TaylorF2_3.5PN")])
def test_smaller_networks_and_phase_marginalisation_on_the_node_axis(c2, n_ifo, marg, approx):
    """the other instantiations of the unified launch: one and two detectors, the generic (K = 28) phase basis, ln I0"""
    cfg, gold, net, names, const = c2
    ifos, datas, dets, noises, f = net
    sub = (ifos[:n_ifo], {k: datas[k] for k in ifos[:n_ifo]}, {k: dets[k] for k in ifos[:n_ifo]},
           {k: noises[k] for k in ifos[:n_ifo]}, f)
    comp = _like(cfg, sub, approx=approx, marg_phi_ref=marg)
    ref = _like(cfg, sub, approx=approx, marg_phi_ref=marg, sincos="fp64")
    assert comp.engine.compress_info(0)["nodes"] > 0
    Xn, _ = syn.draw_walkers(cfg, 1500, 61, "narrow")
    Xw, _ = syn.draw_walkers(cfg, 1500, 62, "wide")
    X = np.concatenate([Xn, Xw])
    X[5::50, names.index("time_shift")] = 0.5          # some on level 1
    X[7::50, names.index("mchirp")] = 0.2              # some on the full grid
    r = _ratio(comp.log_like_batch(X, names, const), ref.log_like_batch(X, names, const))
    counts = comp.engine.last_level_counts()
    print("%d detector(s), marg_phi=%s, %s: %.3f tol, tables %s" % (n_ifo, marg, approx, r, counts))
    assert r <= 0.8 and min(counts) > 0


def test_focus_table_around_a_point(c2):
    """the compressed axis taken to its limit around one point: posterior-scale walkers on a few hundred FP64 nodes"""
    cfg, gold, net, names, const = c2
    like, ref = _like(cfg, net), _like(cfg, net, sincos="fp64")
    inj = syn.injection_params(cfg)
    Xn, _ = syn.draw_walkers(cfg, 3000, 71, "narrow")
    Xw, _ = syn.draw_walkers(cfg, 1000, 72, "wide")
    X = np.concatenate([Xn, Xw])
    before = like.log_like_batch(X, names, const)
    info = like.focus(inj)
    assert 64 <= info["nodes"] <= 1024 and info["kept_bins"] == 0, info
    got = like.log_like_batch(X, names, const)
    c0, c1, cf, cfoc = like.engine.last_level_counts()
    assert c0 + c1 + cf + cfoc == len(X) and cfoc >= 2900, (c0, c1, cf, cfoc)
    want = ref.log_like_batch(X, names, const)
    levels = like.engine.last_levels(len(X))
    in_focus = levels == 3
    # focused walkers: all FP64 on ~300 nodes -- far inside the tolerance; the others took the path they took before
    assert _ratio(got[in_focus], want[in_focus]) <= 0.05
    assert np.array_equal(got[~in_focus], before[~in_focus])
    assert _ratio(got, want) <= 1.0
    # the focus follows the sampler: a fiducial elsewhere leaves these walkers on the ordinary tables, bit for bit
    far = dict(inj)
    far["mchirp"] = inj["mchirp"] * 1.01
    like.focus(far)
    again = like.log_like_batch(X, names, const)
    assert like.engine.last_level_counts()[3] < 50
    lv2 = like.engine.last_levels(len(X)) == 3
    assert np.array_equal(again[~lv2], before[~lv2])
    like.unfocus()
    assert np.array_equal(like.log_like_batch(X, names, const), before) and len(like.engine.last_level_counts()) == 3


def test_focus_with_two_detectors_generic_phase_basis_and_phase_marginalisation(c2):
    cfg, gold, net, names, const = c2
    ifos, datas, dets, noises, f = net
    sub = (ifos[:2], {k: datas[k] for k in ifos[:2]}, {k: dets[k] for k in ifos[:2]}, {k: noises[k] for k in ifos[:2]}, f)
    approx = "TaylorF2_5.5PN_7.5PNTides2020"
    like = _like(cfg, sub, approx=approx, marg_phi_ref=True)
    ref = _like(cfg, sub, approx=approx, marg_phi_ref=True, sincos="fp64")
    X, _ = syn.draw_walkers(cfg, 2000, 81, "narrow")
    # the fiducial as a row of the batch (names / const form of the call)
    info = like.focus(X[0], names=names, const=const)
    assert 64 <= info["nodes"] <= 2048
    got = like.log_like_batch(X, names, const)
    counts = like.engine.last_level_counts()
    assert counts[3] >= 1900, counts
    assert _ratio(got, ref.log_like_batch(X, names, const)) <= 0.3
    # per-detector inner products of a focused point come out of the same epilogue
    p = {k: float(v) for k, v in zip(names, X[1])}
    p.update(const)
    a, b = like.inner_products(dict(p)), ref.inner_products(dict(p))
    for ifo in like.ifos:
        assert abs(a[ifo][0] - b[ifo][0]) <= 1e-6 + 1e-9 * abs(b[ifo][0]) and abs(a[ifo][1] - b[ifo][1]) <= 1e-9 * abs(b[ifo][1])
