"""
Synthetic injections for tests and benchmarks: design-PSD Gaussian noise drawn
directly in the frequency domain with a fixed seed, plus a TaylorF2 signal, wrapped exactly like the
reference's ``--fd-inject`` path (bajes/pipe/gw_init.py:72-85; noise statistics of
bajes/obs/gw/noise.py:211-215).

The builder is class-agnostic: it is handed the Series / Noise / Detector / Waveform classes to use,
so the same recipe drives the product classes (GPU), the oracle port (CPU) and -- in the build
container only -- the unmodified reference.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

T_GPS = 1187008882.0          # default of scripts/bajes_inject:464

INJECTIONS = {
    "bbh": dict(mchirp=30.0, q=1.5, s1z=0.3, s2z=-0.2, lambda1=0.0, lambda2=0.0, distance=400.0,
                cos_iota=0.3, ra=1.2, dec=-0.5, psi=0.7, time_shift=0.003, phi_ref=1.1),
    "bns": dict(mchirp=1.1977, q=1.3, s1z=0.02, s2z=0.01, lambda1=400.0, lambda2=600.0, distance=40.0,
                cos_iota=float(np.cos(2.5)), ra=3.446, dec=-0.408, psi=0.3, time_shift=0.003, phi_ref=1.0),
}

# BASELINE.json configs (C1, C2) and reduced variants used where the CPU oracle has to finish in seconds
CONFIGS = {
    "C1": dict(ifos=["H1", "L1"], seglen=4.0, srate=4096.0, f_min=20.0, f_max=1024.0,
               approx="TaylorF2_3.5PN", source="bbh", n_walkers=64, seed=1001),
    "C2": dict(ifos=["H1", "L1", "V1"], seglen=128.0, srate=8192.0, f_min=20.0, f_max=4096.0,
               approx="TaylorF2_3.5PN", source="bns", n_walkers=4096, seed=2002),
    "C2_small": dict(ifos=["H1", "L1", "V1"], seglen=16.0, srate=4096.0, f_min=20.0, f_max=2048.0,
                     approx="TaylorF2_3.5PN", source="bns", n_walkers=256, seed=2102),
    "C2_55": dict(ifos=["H1", "L1", "V1"], seglen=128.0, srate=8192.0, f_min=20.0, f_max=4096.0,
                  approx="TaylorF2_5.5PN_7.5PNTides", source="bns", n_walkers=4096, seed=2002),
    "C4": dict(ifos=["H1", "L1", "V1"], seglen=256.0, srate=4096.0, f_min=20.0, f_max=2048.0,
               approx="TaylorF2_3.5PN", source="bns", n_walkers=1 << 20, seed=4004),
    "C4_small": dict(ifos=["H1", "L1", "V1"], seglen=32.0, srate=4096.0, f_min=20.0, f_max=2048.0,
                     approx="TaylorF2_3.5PN", source="bns", n_walkers=4096, seed=4104),
}

WALKER_NAMES = ["mchirp", "q", "s1z", "s2z", "lambda1", "lambda2", "distance", "cos_iota",
                "ra", "dec", "psi", "time_shift", "phi_ref"]


def constants(cfg):
    """The constant block bajes puts in the prior (bajes/pipe/gw_init.py:405-408,860-871)."""
    return dict(f_min=cfg["f_min"], f_max=cfg["f_max"], t_gps=T_GPS, seglen=cfg["seglen"], srate=cfg["srate"],
                tukey=0.1, s1x=0.0, s1y=0.0, s2x=0.0, s2y=0.0)       # aligned spins: gw_init.py:405-408


def injection_params(cfg):
    p = dict(INJECTIONS[cfg["source"]])
    p.update(constants(cfg))
    return p


def frequency_axis(cfg):
    n = int(cfg["seglen"] * cfg["srate"] / 2) + 1
    return np.arange(n) / cfg["seglen"]


def draw_noise(psd, seglen, seed):
    """sigma_k = sqrt(S_k T)/2 per quadrature (noise.py:211-215 with df = 1/T)."""
    rng = np.random.default_rng(seed)
    sigma = 0.5 * np.sqrt(psd * seglen)
    re = rng.standard_normal(len(psd))
    im = rng.standard_normal(len(psd))
    with np.errstate(invalid="ignore"):
        d = sigma * re + 1j * sigma * im
    d[~np.isfinite(sigma)] = 0.0
    return d


def draw_walkers(cfg, n, seed, box="narrow"):
    """Batch of parameter points around the injection ('narrow': posterior scale, 'wide': prior scale);
    boxes of SURVEY.md section 8d.  Returns (X [n, 13], names)."""
    rng = np.random.default_rng(seed)
    inj = INJECTIONS[cfg["source"]]
    u = lambda lo, hi: rng.uniform(lo, hi, n)
    if cfg["source"] == "bns":
        if box == "narrow":
            # the likelihood width in chirp mass shrinks ~ 1/seglen-ish: scale the box with the segment
            rel = 5e-6 * (128.0 / cfg["seglen"])
            cols = dict(mchirp=inj["mchirp"] * (1 + u(-rel, rel)), q=np.maximum(1.0, inj["q"] + u(-0.05, 0.05)),
                        s1z=inj["s1z"] + u(-0.01, 0.01), s2z=inj["s2z"] + u(-0.01, 0.01),
                        lambda1=u(0, 1500), lambda2=u(0, 1500), distance=inj["distance"] * (1 + u(-0.2, 0.2)),
                        cos_iota=np.clip(inj["cos_iota"] + u(-0.1, 0.1), -1, 1),
                        ra=inj["ra"] + u(-0.01, 0.01), dec=inj["dec"] + u(-0.01, 0.01), psi=u(0, np.pi),
                        time_shift=inj["time_shift"] + u(-2e-4, 2e-4), phi_ref=u(0, 2 * np.pi))
        else:
            cols = dict(mchirp=u(0.8, 2.0), q=u(1, 2), s1z=u(-0.05, 0.05), s2z=u(-0.05, 0.05),
                        lambda1=u(0, 5000), lambda2=u(0, 5000), distance=u(10, 100), cos_iota=u(-1, 1),
                        ra=u(0, 2 * np.pi), dec=np.arcsin(u(-1, 1)), psi=u(0, np.pi),
                        time_shift=u(-0.05, 0.05), phi_ref=u(0, 2 * np.pi))
    else:
        if box == "narrow":
            cols = dict(mchirp=inj["mchirp"] * (1 + u(-2e-3, 2e-3)), q=np.maximum(1.0, inj["q"] + u(-0.1, 0.1)),
                        s1z=inj["s1z"] + u(-0.05, 0.05), s2z=inj["s2z"] + u(-0.05, 0.05),
                        lambda1=np.zeros(n), lambda2=np.zeros(n), distance=inj["distance"] * (1 + u(-0.2, 0.2)),
                        cos_iota=np.clip(inj["cos_iota"] + u(-0.1, 0.1), -1, 1),
                        ra=inj["ra"] + u(-0.02, 0.02), dec=inj["dec"] + u(-0.02, 0.02), psi=u(0, np.pi),
                        time_shift=inj["time_shift"] + u(-5e-4, 5e-4), phi_ref=u(0, 2 * np.pi))
        else:
            cols = dict(mchirp=u(25, 45), q=u(1, 3), s1z=u(-0.8, 0.8), s2z=u(-0.8, 0.8),
                        lambda1=np.zeros(n), lambda2=np.zeros(n), distance=u(50, 1000), cos_iota=u(-1, 1),
                        ra=u(0, 2 * np.pi), dec=np.arcsin(u(-1, 1)), psi=u(0, np.pi),
                        time_shift=u(-0.05, 0.05), phi_ref=u(0, 2 * np.pi))
    X = np.stack([cols[k] for k in WALKER_NAMES], axis=1)
    return np.ascontiguousarray(X), list(WALKER_NAMES)


def product_classes():
    """The bajes_b200 classes (GPU engine) in the shape ``build_network`` expects."""
    from .detector import Detector
    from .noise import Noise, design_asd
    from .series import Series
    from .waveform import Waveform
    return SimpleNamespace(Series=Series, Noise=Noise, Detector=Detector, Waveform=Waveform, design_asd=design_asd,
                           make_series=lambda d, f, cfg: Series("freq", d, srate=cfg["srate"], seglen=cfg["seglen"],
                                                                f_min=cfg["f_min"], f_max=cfg["f_max"], t_gps=T_GPS,
                                                                only=True, importfreqs=f))


def project_signal(det, hphc, freqs_band, p):
    """F+ hp + Fx hc with the geocentric-delay ramp (detector.py:394-418), host arithmetic for set-up."""
    t = p["t_gps"] + p["time_shift"]
    fp, fc = det.antenna_pattern(p["ra"], p["dec"], p["psi"], t)
    delay = det.time_delay_from_earth_center(p["ra"], p["dec"], t) + p["time_shift"]
    return (fp * hphc.plus + fc * hphc.cross) * np.exp(-2j * np.pi * freqs_band * delay)


def build_network(cfg, classes, inject=True, noise=True, dets=None):
    """ifos, datas, dets, noises, freqs ready for ``GWLikelihood(ifos, datas, dets, noises, freqs, ...)``."""
    f = frequency_axis(cfg)
    band = (f >= cfg["f_min"]) & (f <= cfg["f_max"])
    p = injection_params(cfg)
    datas, noises = {}, {}
    dets = dict(dets) if dets is not None else {}
    hphc = None
    for i, ifo in enumerate(cfg["ifos"]):
        fa, asd = classes.design_asd(ifo)
        noises[ifo] = classes.Noise(fa, asd, f_min=cfg["f_min"], f_max=cfg["f_max"])
        if ifo not in dets:
            dets[ifo] = classes.Detector(ifo, T_GPS)
        psd = noises[ifo].interp_psd_pad(f)
        d = draw_noise(psd, cfg["seglen"], cfg["seed"] + i) if noise else np.zeros(len(f), dtype=complex)
        if inject:
            if hphc is None:
                wave = classes.Waveform(f[band], cfg["srate"], cfg["seglen"], cfg["approx"])
                hphc = wave.compute_hphc(dict(p))
            d[band] += project_signal(dets[ifo], hphc, f[band], p)
        datas[ifo] = classes.make_series(d, f, cfg)
    return cfg["ifos"], datas, dets, noises, f


def make_roq_basis(cfg, n_lin=48, n_quad=24, seed=7):
    """Deterministic stand-in for a JenpyROQ basis on the in-band axis of ``cfg``:
    returns dict(lin_f, quad_f, lin_b [N, n_lin] complex, quad_b [N, n_quad] real).  Node frequencies
    are a subset of the axis (so the joined axis is on-grid); the interpolants are smooth bumps with a
    slowly varying complex phase -- enough to exercise every code path."""
    rng = np.random.default_rng(seed)
    f = frequency_axis(cfg)
    fb = f[(f >= cfg["f_min"]) & (f <= cfg["f_max"])]
    n = len(fb)
    il = np.unique(np.round(np.geomspace(1, n - 2, n_lin)).astype(int))
    iq = np.unique(np.round(np.geomspace(2, n - 3, n_quad)).astype(int))
    x = np.linspace(0.0, 1.0, n)

    def bumps(idx, cplx):
        cols = []
        for i in idx:
            width = 0.01 + 0.05 * rng.random()
            b = np.exp(-0.5 * ((x - x[i]) / width) ** 2)
            if cplx:
                b = b * np.exp(1j * (rng.normal() + 6.0 * rng.random() * (x - x[i])))
            cols.append(b / np.sqrt(n))
        return np.stack(cols, axis=1)

    return dict(lin_f=fb[il], quad_f=fb[iq], lin_b=bumps(il, True), quad_b=np.real(bumps(iq, False)))


def write_jenpyroq(path, basis, cfg):
    """JenpyROQ directory layout read by roq.py:834-853."""
    os.makedirs(os.path.join(path, "ROQ_data/linear"), exist_ok=True)
    os.makedirs(os.path.join(path, "ROQ_data/quadratic"), exist_ok=True)
    np.save(os.path.join(path, "ROQ_data/linear/empirical_frequencies_linear.npy"), basis["lin_f"])
    np.save(os.path.join(path, "ROQ_data/quadratic/empirical_frequencies_quadratic.npy"), basis["quad_f"])
    np.save(os.path.join(path, "ROQ_data/linear/basis_interpolant_linear.npy"), basis["lin_b"])
    np.save(os.path.join(path, "ROQ_data/quadratic/basis_interpolant_quadratic.npy"), basis["quad_b"])
    with open(os.path.join(path, "ROQ_data/ROQ_metadata.txt"), "w") as fh:
        fh.write("# fmin fmax seglen\n%.17g %.17g %.17g\n" % (cfg["f_min"], cfg["f_max"], cfg["seglen"]))



def extras_setup(cfg, nspcal=0, nweights=0):
    """spcal node frequencies and PSD-weight band lengths exactly as bajes builds them
    (bajes/pipe/gw_init.py:690-703,729): log-spaced between the first and last in-band frequency."""
    f = frequency_axis(cfg)
    fb = f[(f >= cfg["f_min"]) & (f <= cfg["f_max"])]
    out = dict(nspcal=nspcal, spcal_freqs=None, nweights=nweights, len_weights=None)
    if nweights:
        f_cut = np.logspace(1.0, np.log(np.max(fb)) / np.log(np.min(fb)), base=np.min(fb), num=nweights + 1)
        lw = np.array([len(fb[np.where((fb >= f_cut[i]) & (fb < f_cut[i + 1]))]) for i in range(nweights)])
        lw[-1] += 1
        lw[-1] += len(fb) - int(np.sum(lw))
        out["len_weights"] = lw
    if nspcal:
        out["spcal_freqs"] = np.logspace(1.0, np.log(np.max(fb)) / np.log(np.min(fb)), base=np.min(fb), num=nspcal)
    return out


def extras_columns(cfg, n, seed, nspcal=0, nweights=0, amp_sigma=0.05, phi_sigma=0.05):
    """Random calibration / weight parameters: (X_extra [n, m], names) with the bajes naming
    (bajes/obs/gw/detector.py:17-23)."""
    rng = np.random.default_rng(seed)
    cols, names = [], []
    for ifo in cfg["ifos"]:
        for j in range(nspcal):
            names.append("spcal_amp{}_{}".format(j, ifo)); cols.append(rng.normal(0.0, amp_sigma, n))
            names.append("spcal_phi{}_{}".format(j, ifo)); cols.append(rng.normal(0.0, phi_sigma, n))
        for b in range(nweights):
            names.append("weight{}_{}".format(b, ifo)); cols.append(rng.uniform(0.7, 1.4, n))
    return (np.stack(cols, axis=1) if cols else np.zeros((n, 0))), names
