"""
Gaussian stationary noise model mirroring ``bajes.obs.gw.Noise`` (bajes/obs/gw/noise.py:95-173).
Set-up only: the likelihood asks it once per detector for the PSD on the in-band axis
(``interp_psd_pad``, noise.py:172-173 -> detector.py:483).
"""
from __future__ import annotations

import os

import numpy as np

_DESIGN_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "design_asd.npz")

# detector -> key of the design curve (noise.py:52-72)
_DESIGN_KEY = {"H1": "aLIGO", "L1": "aLIGO", "I1": "aLIGO", "V1": "AdV"}


def design_asd(ifo: str):
    """Frequency / ASD columns of the design sensitivity curve of ``ifo``
    (``read_asd('design', ifo)``, obs/gw/utils/__init__.py:277-285 -> noise.py:46-75).

    The tables are the public LIGO-P1200087-v18 design curves (aLIGO, AdV); they are stored as a
    compressed fixture, see ``oracle/make_golden.py``."""
    if ifo not in _DESIGN_KEY:
        raise AttributeError("Design ASD not bundled for detector %s (available: %s)" % (ifo, sorted(_DESIGN_KEY)))
    with np.load(_DESIGN_FILE) as z:
        key = _DESIGN_KEY[ifo]
        return np.array(z[key + "_f"]), np.array(z[key + "_asd"])


class Noise(object):
    """ASD table -> linearly interpolated PSD, flat-padded down to f=0 and up to 2 f_max."""

    def __init__(self, freqs, asd, f_min=None, f_max=None, filter=False):
        if len(freqs) != len(asd):
            raise RuntimeError("Numbers of data points do not match. Please check the input ASD data.")
        if filter:
            raise NotImplementedError("band-pass filtered PSDs are outside the likelihood path")
        freqs = np.asarray(freqs, dtype=float)
        asd = np.asarray(asd, dtype=float)
        order = np.argsort(freqs)
        self.freqs = freqs[order]
        self.df = np.median(np.diff(self.freqs))
        self.f_max = np.max(self.freqs) if f_max is None else f_max
        self.f_min = np.min(self.freqs) if f_min is None else f_min
        self.amp_spectrum = asd[order]
        self.power_spectrum = self.amp_spectrum * self.amp_spectrum

        fpad, apad, ppad = self.freqs, self.amp_spectrum, self.power_spectrum
        lo = self.freqs[0]
        if lo > self.df:                                   # noise.py:143-148
            n = int(round(lo / self.df)) + 1
            fpad = np.concatenate([lo - np.arange(n, 0, -1) * self.df, fpad])
            apad = np.concatenate([np.full(n, self.amp_spectrum[0]), apad])
            ppad = np.concatenate([np.full(n, self.power_spectrum[0]), ppad])
        top = fpad[-1]
        if self.f_max * 2 > top:                           # noise.py:150-155
            n = int(round((self.f_max * 2 - top) / self.df)) + 1
            fpad = np.concatenate([fpad, top + np.arange(1, n + 1) * self.df])
            apad = np.concatenate([apad, np.full(n, self.amp_spectrum[-1])])
            ppad = np.concatenate([ppad, np.full(n, self.power_spectrum[-1])])
        self.freqs_pad = fpad
        self.amp_spec_pad = apad
        self.pow_spec_pad = ppad

    @staticmethod
    def _lin_interp(x, y, fr):
        """scipy ``interp1d(kind='linear', bounds_error=False, fill_value=inf)`` (noise.py:157-160)."""
        fr = np.asarray(fr, dtype=float)
        hi = np.clip(np.searchsorted(x, fr), 1, len(x) - 1)
        lo = hi - 1
        slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
        out = slope * (fr - x[lo]) + y[lo]
        return np.where((fr < x[0]) | (fr > x[-1]), np.inf, out)

    def interp_psd_pad(self, fr):
        return self._lin_interp(self.freqs_pad, self.pow_spec_pad, fr)

    def interp_asd_pad(self, fr):
        return self._lin_interp(self.freqs_pad, self.amp_spec_pad, fr)

    def interp_asd(self, fr):
        return self._lin_interp(self.freqs, self.amp_spectrum, fr)

    def interp_psd(self, fr):
        a = self.interp_asd(fr)
        return a * a
