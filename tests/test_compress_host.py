"""
Compressed frequency axis (capi.cu: build_compressed_axis) -- host-side checks, no GPU.

The construction has no counterpart in the reference: it regroups the bin sum of ``Detector.compute_inner_products``
(bajes/obs/gw/detector.py:541) so that ~20x fewer terms reproduce it.  What is checked here:
  * against an independent NumPy construction of the same Chebyshev-node weights;
  * that sum_j W_j g(f_j) = sum_k w_k g(f_k) for smooth g inside the design (to rounding), and that it does NOT hold far
    outside (which is why every call classifies its walkers);
  * against the oracle port: the per-detector <d|h> of real TaylorF2 points, node axis versus every bin.
"""
import numpy as np
import pytest

from bajes_b200 import engine, synthetic as syn
from oracle import bajes_port as port, harness

GT = 4.92549094830932e-6      # taylorf2.py:6


def design_budget(f, mchirp_min=0.8, tau_max=0.11, margin=1.3):
    return margin * 5.0 / 256.0 * (mchirp_min * GT) ** (-5.0 / 3.0) * (np.pi * f) ** (-8.0 / 3.0) + tau_max


def numpy_axis(f, w, n=64, reach=32.0, **kw):
    """independent restatement: greedy runs, first-kind Chebyshev nodes, Lagrange cardinal sums (plain products)"""
    nf, nw = [], []
    k, N = 0, len(f)
    j = np.arange(n)
    xn = -np.cos(np.pi * (2 * j + 1) / (2 * n))
    while k < N:
        width = 2.0 * reach / (2.0 * np.pi * design_budget(f[k], **kw))
        k1 = k + 1
        while k1 < N and f[k1] - f[k] <= width:
            k1 += 1
        if k1 - k <= n + n // 2:
            ke = min(N, k + n)
            nf.append(f[k:ke]); nw.append(w[:, k:ke]); k = ke
            continue
        a, b = f[k], f[k1 - 1]
        x = (f[k:k1] - 0.5 * (a + b)) / (0.5 * (b - a))
        ell = np.ones((k1 - k, n), dtype=np.longdouble)
        for m in range(n):                         # l_m(x) = prod_{q != m} (x - x_q) / (x_m - x_q)
            for q in range(n):
                if q != m:
                    ell[:, m] *= (x.astype(np.longdouble) - xn[q]) / (np.longdouble(xn[m]) - xn[q])
        nf.append(0.5 * (a + b) + 0.5 * (b - a) * xn)
        nw.append((w[:, k:k1].astype(np.clongdouble) @ ell).astype(np.complex128))
        k = k1
    return np.concatenate(nf), np.concatenate(nw, axis=1)


def smooth_g(fr, mchirp, tau):
    psi = 3.0 / 128.0 * (np.pi * mchirp * GT * fr) ** (-5.0 / 3.0)
    return (1.0 + 1e-3 * fr ** (2.0 / 3.0)) * np.exp(-1j * (psi + 2.0 * np.pi * fr * tau))


@pytest.fixture(scope="module")
def axis():
    T = 64.0
    f = np.arange(int(20 * T), int(1024 * T) + 1) / T
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((2, f.size)) + 1j * rng.standard_normal((2, f.size))) * f ** (-7.0 / 6.0)
    return f, w


def test_node_axis_matches_an_independent_numpy_construction(axis):
    f, w = axis
    nf, nw, info = engine.compress_axis(f, w)
    rf, rw = numpy_axis(f, w)
    assert info["nodes"] == len(nf) == len(rf) and info["nodes_per_block"] == 64
    assert np.all(np.diff(nf) > 0), "node frequencies must ascend"
    assert np.allclose(nf, rf, rtol=1e-14, atol=0)
    scale = np.max(np.abs(rw))
    assert np.max(np.abs(nw - rw)) <= 1e-9 * scale      # the plain-product Lagrange form is itself only ~1e-10 accurate
    # the leading plain bins are the grid's own
    kept = 0
    while kept < len(nf) and nf[kept] == f[kept]:
        kept += 1
    assert kept >= 64 and np.array_equal(nw[:, :kept], w[:, :kept])
    assert len(nf) < len(f) / 4


@pytest.mark.parametrize("mchirp,tau", [(0.8, 0.109), (0.9, -0.1), (1.2, 0.003), (30.0, 0.05)])
def test_bin_sum_is_reproduced_inside_the_design(axis, mchirp, tau):
    f, w = axis
    nf, nw, _ = engine.compress_axis(f, w)
    exact = np.sum(w * smooth_g(f, mchirp, tau), axis=1)
    nodes = np.sum(nw * smooth_g(nf, mchirp, tau), axis=1)
    scale = np.sum(np.abs(w * smooth_g(f, mchirp, tau)), axis=1)
    assert np.all(np.abs(exact - nodes) <= 1e-12 * scale)


def test_far_outside_the_design_the_node_sum_is_wrong(axis):
    """|tau| = 1 s against a 0.11 s design: this is what classify_kernel exists for"""
    f, w = axis
    nf, nw, _ = engine.compress_axis(f, w)
    exact = np.sum(w * smooth_g(f, 1.2, 1.0), axis=1)
    nodes = np.sum(nw * smooth_g(nf, 1.2, 1.0), axis=1)
    scale = np.sum(np.abs(w * smooth_g(f, 1.2, 1.0)), axis=1)
    assert np.all(np.abs(exact - nodes) > 1e-8 * scale)
    # and a looser design (level 1 of a context) covers it
    nf1, nw1, info1 = engine.compress_axis(f, w, tau_max=1.1)
    nodes1 = np.sum(nw1 * smooth_g(nf1, 1.2, 1.0), axis=1)
    assert np.all(np.abs(exact - nodes1) <= 1e-12 * scale)
    assert info1["nodes"] > len(nf)


def test_inner_products_of_the_oracle_on_the_node_axis():
    """per-detector <d|h> of TaylorF2 points (oracle port, every bin) versus the same waveform evaluated at the nodes only"""
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=64.0, srate=4096.0)
    like = harness.port_likelihood(cfg)
    ifos, T = like.ifos, cfg["seglen"]
    f = like.dets[ifos[0]].freqs
    w = np.stack([4.0 / T * np.conj(like.dets[i].data) / like.dets[i].psd * np.exp(1j * np.pi * f * T) for i in ifos])
    nf, nw, info = engine.compress_axis(f, w)
    assert info["nodes"] < len(f) / 8
    const = syn.constants(cfg)

    def dh(p, freqs, weights):
        wave = port.WaveformPort(freqs, cfg["srate"], T, cfg["approx"])
        hphc = wave.compute_hphc(dict(p))
        out = []
        for n_, i in enumerate(ifos):
            d = like.dets[i]
            keep = d.freqs
            d.freqs = freqs
            h = d.project_fdwave(hphc, dict(p)) * np.exp(-1j * np.pi * freqs * T)    # centre shift lives in the weights
            d.freqs = keep
            out.append(np.sum(weights[n_] * h))
        return np.array(out)

    Xn, names = syn.draw_walkers(cfg, 3, 5, "narrow")
    Xw, _ = syn.draw_walkers(cfg, 3, 6, "wide")
    for x in np.concatenate([Xn, Xw]):
        p = {k: float(v) for k, v in zip(names, x)}
        p.update(const)
        a, b = dh(p, f, w), dh(p, nf, nw)
        # <d|h> ~ 1e1..1e4 here; the logL tolerance floor is 1e-4
        assert np.max(np.abs(a - b)) <= 1e-9, (p["mchirp"], p["time_shift"], np.abs(a - b))


def test_heterodyned_weights_collapse_the_band_for_walkers_near_a_fiducial():
    """the idea behind the focus table (capi.cu: bj_focus), with the oracle port's waveforms: fold the fiducial waveform's
    phase into the data weights and the factor left over for a nearby walker is so smooth that an axis designed for
    *seconds* of slope reproduces its <d|h> -- while the same axis is useless for the un-heterodyned weights"""
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=64.0, srate=4096.0)
    like = harness.port_likelihood(cfg)
    ifos, T = like.ifos, cfg["seglen"]
    f = like.dets[ifos[0]].freqs
    const = syn.constants(cfg)
    inj = syn.injection_params(cfg)

    def strain(p, freqs):
        wave = port.WaveformPort(freqs, cfg["srate"], T, cfg["approx"])
        hphc = wave.compute_hphc(dict(p))
        out = []
        for i in ifos:
            d = like.dets[i]
            keep = d.freqs
            d.freqs = freqs
            out.append(d.project_fdwave(hphc, dict(p)) * np.exp(-1j * np.pi * freqs * T))
            d.freqs = keep
        return np.array(out)

    w = np.stack([4.0 / T * np.conj(like.dets[i].data) / like.dets[i].psd * np.exp(1j * np.pi * f * T) for i in ifos])
    h_fid = strain(inj, f)
    w_het = w * h_fid / np.abs(h_fid)                        # weights carry the fiducial's phase (per detector)
    # an axis for 0.3 s of slope at f_min falling as the Newtonian chirp, plus 2 ms: a heavy-binary design used as a stand-in
    # for the focus budget a (f/f_lo)^(-5/3) + b
    kw = dict(mchirp_min=12.0, tau_max=2e-3)
    nf, nw_het, info = engine.compress_axis(f, w_het, **kw)
    _, nw_raw, _ = engine.compress_axis(f, w, **kw)
    assert info["nodes"] < len(f) / 100 and info["kept_bins"] == 0, info
    hn_fid = strain(inj, nf)
    X, names = syn.draw_walkers(cfg, 4, 5, "narrow")
    for x in X:
        p = {k: float(v) for k, v in zip(names, x)}
        p.update(const)
        exact = np.sum(w * strain(p, f), axis=1)
        hn = strain(p, nf)
        het = np.sum(nw_het * hn * np.abs(hn_fid) / hn_fid, axis=1)      # walker relative to the fiducial, at the nodes
        raw = np.sum(nw_raw * hn, axis=1)
        assert np.max(np.abs(het - exact)) <= 1e-8, np.abs(het - exact)
        assert np.max(np.abs(raw - exact)) > 1.0                          # without the heterodyne the axis is far too sparse
