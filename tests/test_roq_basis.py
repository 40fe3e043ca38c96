"""
ROQ basis construction (bajes_b200/roq_basis.py, csrc/roq_basis.cuh).  The reference's own
construction is commented-out code (bajes/pipe/utils/roq.py:16-662): parity is UNPINNED; the checker is the NumPy
restatement of that code in oracle/roq_basis_port.py plus the properties the construction must have.

CPU: the restatement (orthonormality under the real inner product, greedy errors fall below epsilon, DEIM nodes distinct,
     interpolation of the training set) and the product's host-side DEIM against it.
GPU: training bank vs the oracle port's TaylorF2; greedy picks / basis vs the restatement; a basis built on the GPU drives the
     ROQ likelihood kernel and reproduces the full-grid likelihood.
"""
import numpy as np
import pytest

from bajes_b200 import synthetic as syn
from bajes_b200.roq_basis import deim, draw_training_parameters
from oracle import bajes_port as port, roq_basis_port as rb

NAMES = ["mchirp", "q", "s1z", "s2z", "lambda1", "lambda2", "distance", "cos_iota", "phi_ref"]
RANGES = {"mchirp": (1.19, 1.21), "q": (1.0, 1.4), "s1z": (-0.02, 0.02), "s2z": (-0.02, 0.02), "lambda1": (0.0, 800.0),
          "lambda2": (0.0, 800.0), "distance": 100.0, "cos_iota": 1.0, "phi_ref": 0.0}
CFG = dict(seglen=2.0, srate=2048.0, f_min=30.0, f_max=512.0, approx="TaylorF2_3.5PN")


def _axis():
    f = np.arange(int(CFG["f_min"] * CFG["seglen"]), int(CFG["f_max"] * CFG["seglen"]) + 1) / CFG["seglen"]
    fa, asd = syn.product_classes().design_asd("H1")
    psd = np.interp(f, fa, asd) ** 2
    return f, psd


def _port_bank(f, X, quadratic=False):
    wave = port.WaveformPort(f, CFG["srate"], CFG["seglen"], CFG["approx"])
    rows = []
    for x in X:
        p = {k: float(v) for k, v in zip(NAMES, x)}
        p.update(f_min=CFG["f_min"], f_max=CFG["f_max"], seglen=CFG["seglen"], srate=CFG["srate"], s1x=0.0, s1y=0.0, s2x=0.0,
                 s2y=0.0, tukey=0.1, t_gps=syn.T_GPS)
        hp = wave.compute_hphc(p, roq=True).plus          # roq: no centre shift (waveform.py:390)
        rows.append(np.abs(hp) ** 2 + 0j if quadratic else hp)
    return np.array(rows)


def test_restated_greedy_and_deim_properties():
    f, psd = _axis()
    w, df = 1.0 / psd, 1.0 / CFG["seglen"]
    X = draw_training_parameters(RANGES, 40, np.random.default_rng(2), NAMES)
    assert len(X) == 2 ** 6 + 40 and np.all(X[:64, 0] != 1.2)           # corners first
    T = np.array([rb.normalize(w, h, df) for h in _port_bank(f, X)])
    basis, picks, sig = rb.rb_greedy(T, w, df, 1e-5)
    m = len(basis)
    assert picks[0] == 0 and len(set(picks)) == m and 3 <= m < len(T)
    assert sig[-1] <= 1e-5 and np.all(np.diff(sig[1:]) < 1e-12 + 0.5 * sig[1:-1])  # errors fall (not strictly monotone)
    G = np.array([[rb.inner_product(w, a, b, df) for b in basis] for a in basis])
    assert np.max(np.abs(G - np.eye(m))) < 1e-4      # classical Gram-Schmidt, as the reference runs it: orthogonal to ~sigma_end
    V = np.array(basis).T
    nodes, p = rb.deim(V, f)
    assert len(set(p.tolist())) == m and np.array_equal(nodes, f[p])
    # the product's host-side DEIM is the same algorithm
    nodes2, p2 = deim(V, f)
    assert np.array_equal(p, p2) and np.array_equal(nodes, nodes2)
    B = rb.interpolant(V, p)
    assert np.allclose(B[p, :], np.eye(m), atol=1e-8)
    # every training function is reproduced from its values at the nodes
    err = max(rb.norm(w, T[i] - B @ T[i][p], df) for i in range(0, len(T), 7))
    assert err < 1e-2          # greedy error x Lebesgue constant of the nodes


def test_real_inner_product_of_the_commented_code_breaks_deim():
    """why the restatement repairs roq.py:662: with Re(a, b) the basis holds C-dependent vectors and V[p, :] is singular"""
    f, psd = _axis()
    w, df = 1.0 / psd, 1.0 / CFG["seglen"]
    X = draw_training_parameters(RANGES, 40, np.random.default_rng(2), NAMES)
    T0 = _port_bank(f, X)
    conds = {}
    for herm in (False, True):
        rb.HERMITIAN = herm
        try:
            T = np.array([rb.normalize(w, h, df) for h in T0])
            basis, picks, sig = rb.rb_greedy(T, w, df, 1e-5)
            V = np.array(basis).T
            _, p = rb.deim(V, f)
            conds[herm] = (len(basis), np.linalg.cond(V[p, :]))
        finally:
            rb.HERMITIAN = True
    assert conds[False][0] > 1.5 * conds[True][0] and conds[False][1] > 1e6 * conds[True][1], conds


@pytest.mark.gpu
def test_training_bank_and_greedy_on_the_gpu_against_the_restatement():
    import torch
    from bajes_b200 import roq_basis as rbg
    f, psd = _axis()
    w, df = 1.0 / psd, 1.0 / CFG["seglen"]
    X = draw_training_parameters(RANGES, 60, np.random.default_rng(3), NAMES)
    for quad in (False, True):
        bank = rbg._bank(0, CFG["approx"], f, X, NAMES, {}, quad)
        ref = _port_bank(f, X, quad)
        got = bank.cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-9 * np.max(np.abs(ref))
    # greedy: same picks, same basis
    bank = rbg._bank(0, CFG["approx"], f, X, NAMES, {}, False)
    T = np.array([rb.normalize(w, h, df) for h in bank.cpu().numpy()])
    want_basis, want_picks, want_sig = rb.rb_greedy(T, w, df, 1e-6)
    train = torch.as_tensor(T, device="cuda:0").contiguous()
    basis = torch.zeros((256, len(f)), dtype=torch.complex128, device="cuda:0")
    n, picks, sig, ms = rbg._greedy(0, train, w, df, 1e-6, basis, 0, 256)
    # identical picks while the errors are far above the rounding of the two orthogonalisations (the restatement keeps
    # projections of the ORIGINAL functions as the reference does, the device kernel subtracts from the residuals)
    k = int(np.sum(want_sig[:-1] > 1e-4))
    assert k >= 5 and picks[:k].tolist() == want_picks[:k]
    assert np.allclose(sig[:k], want_sig[:k], rtol=1e-5)
    assert np.max(np.abs(basis[:k].cpu().numpy() - np.array(want_basis[:k]))) < 1e-5
    assert abs(n - len(want_basis)) <= 3 and sig[-1] <= 1e-6
    # the device basis is orthonormal under the real inner product, and spans the training set to the tolerance
    Bn = basis[:n].cpu().numpy()
    G = 4.0 * df * ((np.conj(Bn) * w) @ Bn.T)
    assert np.max(np.abs(G - np.eye(n))) < 1e-8
    worst = max(rb.norm(w, T[i] - rb.project(T[i], Bn, w, df), df) for i in range(len(T)))
    assert worst <= 2e-6
    # continuing from a prebuilt basis (roq.py:512-520) adds nothing when the new training set is the old one
    train2 = torch.as_tensor(T, device="cuda:0").contiguous()
    n2, picks2, sig2, _ = rbg._greedy(0, train2, w, df, 2e-6, basis, n, 256)
    assert n2 == n and len(picks2) == 0 and sig2[-1] <= 2e-6


@pytest.mark.gpu
def test_a_basis_built_on_the_gpu_reproduces_the_full_likelihood():
    """a genuine small basis -> ROQ logL ~ full logL on points inside the training box"""
    import tempfile
    from bajes_b200 import GWLikelihood, roq as roqmod
    from bajes_b200.ensemble import BoxPrior
    from bajes_b200.roq_basis import build_roq_basis
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=4.0, srate=2048.0, f_min=30.0, f_max=1024.0, seed=3103)
    ifos, datas, dets, noises, f = syn.build_network(cfg, syn.product_classes())
    fb = dets[ifos[0]].freqs if hasattr(dets[ifos[0]], "freqs") and dets[ifos[0]].freqs is not None else None
    like_full = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"])
    fb = like_full.dets[ifos[0]].freqs
    inj = syn.injection_params(cfg)
    ranges = {"mchirp": (inj["mchirp"] * 0.9995, inj["mchirp"] * 1.0005), "q": (1.1, 1.5), "s1z": (-0.03, 0.05),
              "s2z": (-0.03, 0.05), "lambda1": (0.0, 1500.0), "lambda2": (0.0, 1500.0), "distance": 100.0, "cos_iota": 1.0,
              "phi_ref": 0.0}
    psd = np.mean([like_full.dets[i].psd for i in ifos], axis=0)
    basis = build_roq_basis(fb, psd, cfg["seglen"], cfg["approx"], ranges, NAMES, stages=(1500, 1500), tol=1e-7, seed=4)
    n_lin, n_quad = basis["lin_b"].shape[1], basis["quad_b"].shape[1]
    print("basis: %d linear / %d quadratic nodes for %d bins; greedy %.1f / %.1f ms"
          % (n_lin, n_quad, len(fb), basis["lin_info"]["greedy_ms"], basis["quad_info"]["greedy_ms"]))
    assert 8 <= n_lin < 400 and 4 <= n_quad < 200
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    prior = BoxPrior(names, [[-1, 1]] * len(names), {})
    prior.bounds[names.index("time_shift")] = [-0.01, 0.01]
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, basis, cfg)
        roq = roqmod.build_roq_dict(tmp, 2000, ifos, datas, noises, f, cfg["f_min"], cfg["f_max"], cfg["seglen"], prior,
                                    cfg["approx"], device="cuda:0")
    net2 = syn.build_network(cfg, syn.product_classes())
    like_roq = GWLikelihood(*net2, cfg["srate"], cfg["seglen"], cfg["approx"], roq=roq)
    rng = np.random.default_rng(9)
    X = np.tile(np.array([inj[k] for k in names]), (64, 1))
    for k in ("mchirp", "q", "s1z", "s2z", "lambda1", "lambda2"):
        lo, hi = ranges[k]
        X[:, names.index(k)] = rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), 64)
    X[:, names.index("time_shift")] = inj["time_shift"] + rng.uniform(-2e-3, 2e-3, 64)
    a = like_roq.log_like_batch(X, names, const)
    b = like_full.log_like_batch(X, names, const)
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
    d = np.abs(a - b)
    print("ROQ (GPU-built basis) vs full grid: max |dlogL| = %.3g over logL in [%.1f, %.1f]" % (d.max(), b.min(), b.max()))
    assert d.max() < 0.05
