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


def reference_likelihood(cfg, approx=None, marg_phi_ref=False, roq=None, net=None, extras=None):
    from oracle import shims
    shims.import_reference()
    from bajes.pipe.log_like import GWLikelihood
    cl = reference_classes()
    ifos, datas, dets, noises, f = net if net is not None else syn.build_network(cfg, cl)
    return GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], approx or cfg["approx"],
                        marg_phi_ref=marg_phi_ref, roq=roq, **(extras or {}))


def port_likelihood(cfg, approx=None, marg_phi_ref=False, roq=None, net=None, extras=None, **kw):
    cl = port_classes(**kw)
    ifos, datas, dets, noises, f = net if net is not None else syn.build_network(cfg, cl)
    return port.GWLikelihoodPort(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], approx or cfg["approx"],
                                 marg_phi_ref=marg_phi_ref, roq=roq, **(extras or {}))


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


# ---- synthetic ROQ bases: generator lives in bajes_b200/synthetic.py (shared with the benchmarks) -----------
make_roq_basis = syn.make_roq_basis


write_jenpyroq = syn.write_jenpyroq


class FakePrior:
    """The three attributes initialize_roq reads from a bajes Prior (pipe/__init__.py:678-687)."""

    def __init__(self, names, bounds, const):
        self.names = list(names)
        self.bounds = [list(b) for b in bounds]
        self.const = dict(const)
