"""
Strain container mirroring ``bajes.obs.gw.Series`` for the set-up side of the likelihood
(bajes/obs/gw/strain.py:214-379).  Set-up only: it provides ``freqs``, ``freq_series``,
``mask`` and ``window_factor`` to ``Detector.store_measurement``; nothing here is on the
per-point hot path.
"""
from __future__ import annotations

import numpy as np


def freq_axis(n_time: int, dt: float) -> np.ndarray:
    """One-sided FFT axis (strain.py:72-86 with fmin=0)."""
    return np.fft.rfftfreq(n_time, d=dt)


def time_axis(n: int, srate: float, t_gps: float = 0.0) -> np.ndarray:
    """Time stamps centred on ``t_gps`` (strain.py:88-100)."""
    dt = 1.0 / srate
    return np.arange(n) * dt + t_gps - n * dt / 2.0


def tukey_window(h: np.ndarray, alpha: float):
    """Tukey taper and its mean-square ("window factor"), strain.py:145-153."""
    from scipy.signal.windows import tukey
    w = tukey(len(h), alpha)
    return h * w, float(np.mean(w ** 2))


def centre_pad(h: np.ndarray, padlen: int) -> np.ndarray:
    """Symmetric zero padding (strain.py:111-143, where='center')."""
    left = padlen // 2
    right = padlen - left
    return np.concatenate([np.zeros(left), h, np.zeros(right)])


class Series(object):
    """A strain series in the time or frequency domain.

    Constructor signature follows strain.py:218-230.  Supported: ``type='freq'`` (the
    ``--fd-inject`` path, gw_init.py:72-85) and ``type='time'`` with an explicit ``seglen``
    (windowing + one-sided FFT scaled by dt, strain.py:30-52).  Butterworth filtering is not part
    of the likelihood path and raises.
    """

    def __init__(self, type, series, srate, seglen=None, f_min=None, f_max=None, t_gps=0.0,
                 only=False, filter=False, alpha_taper=None, importfreqs=()):
        if filter:
            raise NotImplementedError("Butterworth filtering of the input strain is outside the likelihood path")
        self.window_factor = 1.0
        self.srate = srate
        self.dt = 1.0 / srate
        self.f_Nyq = srate / 2.0
        self.t_gps = t_gps
        self.f_min = 0.0 if f_min is None else f_min
        self.f_max = self.f_Nyq if f_max is None else f_max

        if type == "time":
            series = np.asarray(series, dtype=float)
            if seglen is None:
                raise NotImplementedError("time series without seglen (power-of-two padding) is not mirrored")
            self.seglen = seglen
            n_final = int(np.ceil(seglen * srate))
            if n_final % 2 == 1:
                n_final += 1
            self.alpha_taper = 0.4 / seglen if alpha_taper is None else alpha_taper
            n_raw = len(series)
            if n_raw == n_final:
                self.time_series, wfact = tukey_window(series, self.alpha_taper)
            elif n_final > n_raw:
                from scipy.signal.windows import tukey
                tapered, _ = tukey_window(series, self.alpha_taper)
                pad = n_final - n_raw
                wfact = float(np.mean(np.append(tukey(n_raw, self.alpha_taper) ** 2.0, np.zeros(pad))))
                self.time_series = centre_pad(tapered, pad)
            else:
                half = n_raw // 2
                self.time_series, wfact = tukey_window(series[half - n_final // 2: half + n_final // 2],
                                                       self.alpha_taper)
            self.window_factor = wfact
            self.df = 1.0 / self.seglen
            self.times = time_axis(len(self.time_series), srate, t_gps)
            if only:
                self.freqs, self.freq_series = None, None
            else:
                self.freqs = freq_axis(len(self.time_series), self.dt)
                self.freq_series = np.fft.rfft(self.time_series) * self.dt
        elif type == "freq":
            series = np.asarray(series)
            self.seglen = self.f_Nyq / len(series) if seglen is None else seglen
            self.df = 1.0 / self.seglen
            if len(importfreqs) == 0:
                f = freq_axis((len(series) - 1) * 2, self.dt)
                self.freqs = f[f >= self.f_min] if self.f_min != 0.0 else f
            else:
                self.freqs = np.asarray(importfreqs)
            self.freq_series = series
            assert len(self.freqs) == len(self.freq_series)
            if only:
                self.times, self.time_series = None, None
            else:
                n = int(srate * self.seglen)
                self.time_series = np.fft.irfft(self.freq_series, n) * srate
                self.times = time_axis(n, srate, t_gps)
        else:
            raise AttributeError("Type of series not specified or wrong. Please use 'time' or 'freq'.")

        self.freqs = np.asarray(self.freqs) if self.freqs is not None else None
        if self.freqs is not None:
            # inclusive band (strain.py:375-376)
            self.mask = np.where((self.freqs >= self.f_min) & (self.freqs <= self.f_max))
        else:
            self.mask = None
