"""
TEST INFRASTRUCTURE.  Builders that assemble the SAME synthetic injection (bajes_b200/synthetic.py)
with three interchangeable class sets:

  * ``reference_classes()`` -- the unmodified reference under oracle/shims.py (build container only)
  * ``port_classes()``      -- the NumPy restatement in oracle/bajes_port.py (travels to the GPU box)
  * ``bajes_b200.synthetic.product_classes()`` -- the product (GPU)

and small helpers to evaluate the oracle over a batch.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from bajes_b200 import synthetic as syn          # noqa: E402
from bajes_b200.gmst import SIDEREAL_DAY_S, gmst_from_gps   # noqa: E402
from bajes_b200.noise import design_asd          # noqa: E402
from oracle import bajes_port as port            # noqa: E402


def reference_classes():
    from oracle import shims
    shims.import_reference()
    from bajes.obs.gw import Detector, Noise, Series, Waveform
    from bajes.obs.gw.utils import read_asd

    def make_series(d, f, cfg):
        return Series("freq", d, srate=cfg["srate"], seglen=cfg["seglen"], f_min=cfg["f_min"], f_max=cfg["f_max"],
                      t_gps=syn.T_GPS, only=True, importfreqs=f)

    return SimpleNamespace(Series=Series, Noise=Noise, Detector=Detector, Waveform=Waveform,
                           design_asd=lambda ifo: read_asd("design", ifo), make_series=make_series)


def port_classes(gmst_reference=None, sday=None):
    gm = gmst_from_gps(syn.T_GPS) if gmst_reference is None else gmst_reference
    sd = SIDEREAL_DAY_S if sday is None else sday

    def make_series(d, f, cfg):
        return port.SeriesPort(d, f, cfg["srate"], cfg["seglen"], cfg["f_min"], cfg["f_max"], syn.T_GPS)

    return SimpleNamespace(Series=port.SeriesPort, Noise=port.NoisePort,
                           Detector=lambda ifo, t_gps: port.DetectorPort(ifo, t_gps, gm, sd),
                           Waveform=port.WaveformPort, design_asd=design_asd, make_series=make_series)


def reference_likelihood(cfg, approx=None, marg_phi_ref=False, roq=None, net=None):
    from oracle import shims
    shims.import_reference()
    from bajes.pipe.log_like import GWLikelihood
    cl = reference_classes()
    ifos, datas, dets, noises, f = net if net is not None else syn.build_network(cfg, cl)
    return GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], approx or cfg["approx"],
                        marg_phi_ref=marg_phi_ref, roq=roq)


def port_likelihood(cfg, approx=None, marg_phi_ref=False, roq=None, net=None, **kw):
    cl = port_classes(**kw)
    ifos, datas, dets, noises, f = net if net is not None else syn.build_network(cfg, cl)
    return port.GWLikelihoodPort(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], approx or cfg["approx"],
                                 marg_phi_ref=marg_phi_ref, roq=roq)


def eval_batch(like, X, names, const):
    """[like.log_like({**row, **const}) for row in X] -- the semantics log_like_batch must reproduce."""
    out = np.empty(len(X))
    for r, x in enumerate(X):
        p = {k: float(v) for k, v in zip(names, x)}
        p.update(const)
        out[r] = like.log_like(p)
    return out


def tolerance(ref):
    """|dlogL| <= 1e-6 |logL| + 1e-4  (BASELINE.json north_star)."""
    ref = np.asarray(ref, dtype=float)
    return 1e-6 * np.abs(np.where(np.isfinite(ref), ref, 0.0)) + 1e-4


# ---- synthetic ROQ bases (format-only: smooth random interpolants in the JenpyROQ layout) -------------
def make_roq_basis(cfg, n_lin=48, n_quad=24, seed=7):
    """Deterministic stand-in for a JenpyROQ basis on the in-band axis of ``cfg``:
    returns dict(lin_f, quad_f, lin_b [N, n_lin] complex, quad_b [N, n_quad] real).  Node frequencies
    are a subset of the axis (so the joined axis is on-grid); the interpolants are smooth bumps with a
    slowly varying complex phase -- enough to exercise every code path."""
    rng = np.random.default_rng(seed)
    f = syn.frequency_axis(cfg)
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


class FakePrior:
    """The three attributes initialize_roq reads from a bajes Prior (pipe/__init__.py:678-687)."""

    def __init__(self, names, bounds, const):
        self.names = list(names)
        self.bounds = [list(b) for b in bounds]
        self.const = dict(const)
