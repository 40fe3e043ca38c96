"""
Waveform generator mirroring ``bajes.obs.gw.Waveform`` (bajes/obs/gw/waveform.py:248-395) for the
TaylorF2 family.  ``compute_hphc`` evaluates the polarisations in the CUDA library
(``bj_polarizations``); inside the likelihood the waveform is never materialised at all -- it is
generated and consumed in registers by the fused kernel.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from . import engine as _engine

PolarizationTuple = namedtuple("PolarizationTuple", ("plus", "cross"), defaults=([None], [None]))

# the frequency-domain TaylorF2 entries of waveform.py:25-130
__approx_dict__ = {name: {"type": "fnc", "domain": "freq"} for name in _engine.APPROX_IDS}


def sph2cart(r, theta, phi):
    """bajes/pipe/__init__.py:143-149."""
    return r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)


class Waveform(object):
    """Waveform model object for compact binary coalescences (TaylorF2 family on the GPU)."""

    def __init__(self, freqs, srate, seglen, approx, device=0):
        if approx not in _engine.APPROX_IDS:
            raise AttributeError(
                "Unable to read approximant string '%s'. The B200 likelihood path provides: %s."
                % (approx, list(_engine.APPROX_IDS)))
        self.freqs = np.asarray(freqs, dtype=float)
        self.seglen = seglen
        self.srate = srate
        self.f_min = np.amin(self.freqs)
        self.f_max = np.amax(self.freqs)
        self.df = 1.0 / self.seglen
        self.dt = 1.0 / self.srate
        self.approx = approx
        self.domain = "freq"
        self.device = device

    def compute_hphc(self, params, roq=None):
        """hp, hc on ``self.freqs`` (waveform.py:270-395).  Like the reference it adds the derived
        entries (mtot/mchirp, cartesian spins, iota/cos_iota) to ``params`` and returns an empty
        ``PolarizationTuple`` for unphysical points."""
        if "eos_logp1" in params:
            raise NotImplementedError("EOS-sampled tides (TOV solve) are outside the fused GPU path")
        if "mchirp" in params:
            nu = params["q"] / (1.0 + params["q"]) ** 2.0
            params["mtot"] = params["mchirp"] / nu ** 0.6
        elif "mtot" in params:
            nu = params["q"] / (1.0 + params["q"]) ** 2.0
            params["mchirp"] = params["mtot"] * nu ** 0.6
        if "tilt1" in params and "s1" in params and "phi_1l" in params:
            params["s1x"], params["s1y"], params["s1z"] = sph2cart(params["s1"], params["tilt1"], params["phi_1l"])
        if "tilt2" in params and "s2" in params and "phi_2l" in params:
            params["s2x"], params["s2y"], params["s2z"] = sph2cart(params["s2"], params["tilt2"], params["phi_2l"])
        if "cos_iota" in params:
            params["iota"] = np.arccos(params["cos_iota"])
        elif "iota" in params:
            params["cos_iota"] = np.cos(params["iota"])

        # hand the ORIGINAL parametrisation to the device prologue (it redoes the derivations in FP64)
        sub = {k: v for k, v in params.items() if k in _engine.SLOT_INDEX}
        if "mchirp" in sub and "mtot" in sub:
            del sub["mtot"]
        if "cos_iota" in sub and "iota" in sub:
            del sub["iota"]
        if "s1" in sub and "tilt1" in sub and "phi_1l" in sub:
            for k in ("s1x", "s1y", "s1z"):
                sub.pop(k, None)
        if "s2" in sub and "tilt2" in sub and "phi_2l" in sub:
            for k in ("s2x", "s2y", "s2z"):
                sub.pop(k, None)
        pmap, x = _engine.ParamMap.from_dict(sub)
        hp, hc = _engine.polarizations(self.approx, self.seglen, self.freqs, x, pmap,
                                       centre_shift=(roq is None), device=self.device)
        if np.any(np.isnan(hp)) or not np.any(hp):
            return PolarizationTuple()
        return PolarizationTuple(plus=hp, cross=hc)
