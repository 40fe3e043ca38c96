"""
Ground-based interferometer mirroring ``bajes.obs.gw.Detector`` (bajes/obs/gw/detector.py:164-544).

Host side = geometry constants and measurement storage (set-up).  Everything evaluated per
parameter point -- antenna patterns, geocentric delay, projection, inner products -- runs in the CUDA
library; the per-point methods kept here (``antenna_pattern``, ``time_delay_from_earth_center``,
``project_fdwave``, ``compute_inner_products``) are thin calls into it so that helpers which use them
directly (SNR extraction, injections; bajes/pipe/utils/__init__.py:217-403) keep working.
"""
from __future__ import annotations

import numpy as np

from . import gmst as _gmst

CLIGHT_SI = 299792458.0

# latitude, longitude, elevation, xarm_azimuth, yarm_azimuth, xarm_tilt, yarm_tilt (detector.py:62-160)
_SITES = {
    "H1": (0.81079526383, -2.08405676917, 142.554, 2.1991043855, 3.7699007123, -6.195e-4, 1.25e-5),
    "L1": (0.53342313506, -1.58430937078, -6.574, 3.4508039105, 5.0216002373, 0.0, 0.0),
    "V1": (0.76151183984, 0.18333805213, 51.884, 1.2316334746, 2.8024298014, 0.0, 0.0),
    "K1": (0.6355068497, 2.396441015, 414.181, 0.5166831721, 2.0874762035, 0.0, 0.0),
    "G1": (0.91184982752, 0.17116780435, 114.425, 2.02358884, 0.377195322, 0.0, 0.0),
    "I1": (0.2484185302005, 1.334013324941, 0.0, 2.1991043855, 3.7699007123, 0.0, 0.0),
    "ET1": (0.76151183984, 0.18333805213, 51.884, 0.33916285222, 5.57515060820, 0.0, 0.0),
    "ET2": (0.76299307990, 0.18405858870, 59.735, 4.52795305701, 3.48075550581, 0.0, 0.0),
    "ET3": (0.76270463257, 0.18192996730, 59.727, 2.43355795462, 1.38636040342, 0.0, 0.0),
    "CE": (0.81079526383, -2.08405676917, 142.554, 2.1991043855, 3.7699007123, -6.195e-4, 1.25e-5),
}


def get_detector_information(ifo):
    """latitude, longitude, elevation, xarm_azimuth, yarm_azimuth, xarm_tilt, yarm_tilt."""
    if ifo in _SITES:
        return _SITES[ifo]
    if "CE-North" in ifo:
        return (0.764918, -1.9691740, 0.0, -1.57079632679, 0.0, 0.0, 0.0)
    if "CE-South" in ifo:
        return (-0.593412, 2.5307270, 0.0, -0.7853981634, 0.7853981634, 0.0, 0.0)
    raise ValueError("Can't get information from input detector. Please check you use a correct name. "
                     "Available options: %s." % sorted(_SITES))


def gmst_accurate(gps_time):
    """astropy when installed (as the reference, detector.py:11-15), else the closed form in gmst.py."""
    try:
        from astropy.time import Time
        return Time(gps_time, format="gps", scale="utc", location=(0, 0)).sidereal_time("mean").rad
    except ImportError:
        return _gmst.gmst_from_gps(gps_time)


class Detector(object):

    def __init__(self, detector_init, t_gps=None):
        if isinstance(detector_init, str):
            self.ifo = detector_init
            (self.latitude, self.longitude, self.elevation, self.xarm_azimuth, self.yarm_azimuth,
             self.xarm_tilt, self.yarm_tilt) = get_detector_information(self.ifo)
        elif isinstance(detector_init, list):
            self.ifo = "UNKNOWN"
            (self.latitude, self.longitude, self.elevation, self.xarm_azimuth, self.yarm_azimuth,
             self.xarm_tilt, self.yarm_tilt) = detector_init
        elif isinstance(detector_init, dict):
            self.ifo = "UNKNOWN"
            for k, v in detector_init.items():
                self.__dict__[k] = v
        else:
            raise RuntimeError("Unable to read initialization argument for Detector object. Please use a "
                               "compatible string, a list or a dictionary with the required information.")
        self.x_arm = self.compute_arm(self.xarm_tilt, self.xarm_azimuth)
        self.y_arm = self.compute_arm(self.yarm_tilt, self.yarm_azimuth)
        self.location = self.compute_location()
        self.response = self.compute_response()
        self.reference_time = t_gps
        self.sday = None
        self.gmst_reference = None
        self.compute_gmst_reference()
        self._owner = None           # GWLikelihood that stored a measurement here

    # ---- geometry (set-up; detector.py:207-233) ---------------------------------------------------
    def compute_response(self):
        return 0.5 * (np.outer(self.x_arm, self.x_arm) - np.outer(self.y_arm, self.y_arm))

    def compute_location(self):
        a, b = 6378137, 6356752.314          # WGS-84 semi axes [m]
        sl, cl = np.sin(self.latitude), np.cos(self.latitude)
        r = a ** 2 * (a ** 2 * cl ** 2 + b ** 2 * sl ** 2) ** (-0.5)
        return np.array([(r + self.elevation) * cl * np.cos(self.longitude),
                         (r + self.elevation) * cl * np.sin(self.longitude),
                         ((b / a) ** 2 * r + self.elevation) * sl])

    def compute_arm(self, arm_tilt, arm_azimuth):
        sl, cl = np.sin(self.latitude), np.cos(self.latitude)
        so, co = np.sin(self.longitude), np.cos(self.longitude)
        east = np.array([-so, co, 0])
        north = np.array([-sl * co, -sl * so, cl])
        up = np.array([cl * co, cl * so, sl])
        return (np.cos(arm_tilt) * np.cos(arm_azimuth) * east + np.cos(arm_tilt) * np.sin(arm_azimuth) * north
                + np.sin(arm_tilt) * up)

    def compute_gmst_reference(self):
        if self.reference_time is None:
            raise RuntimeError("Can't get accurate sidereal time without GPS reference time!")
        try:
            from astropy.units.si import sday
            self.sday = float(sday.si.scale)
        except ImportError:
            self.sday = _gmst.SIDEREAL_DAY_S
        self.gmst_reference = float(gmst_accurate(self.reference_time))

    def gmst_estimate(self, gps_time):
        """detector.py:246-253 (scalar set-up helper)."""
        dphase = (gps_time - self.reference_time) / self.sday * (2.0 * np.pi)
        return (self.gmst_reference + dphase) % (2.0 * np.pi)

    def time_delay_from_location(self, other_location, right_ascension, declination, t_gps):
        """``t_here - t_there`` for a signal from (ra, dec) (detector.py:334-357); scalar geometry helper."""
        ang = self.gmst_estimate(t_gps) - right_ascension
        cd = np.cos(declination)
        ehat = np.array([cd * np.cos(ang), cd * -np.sin(ang), np.sin(declination)])
        return (np.asarray(other_location, dtype=float) - self.location).dot(ehat) / CLIGHT_SI

    def time_delay_from_earth_center(self, right_ascension, declination, t_gps):
        """detector.py:318-332; the batched path computes this on the device."""
        return self.time_delay_from_location(np.zeros(3), right_ascension, declination, t_gps)

    def time_delay_from_detector(self, other_detector, right_ascension, declination, t_gps):
        """detector.py:359-376."""
        return self.time_delay_from_location(other_detector.location, right_ascension, declination, t_gps)

    def optimal_orientation(self, t_gps):
        """(ra, dec) overhead of the detector at ``t_gps`` (detector.py:378-392)."""
        return self.longitude + (self.gmst_estimate(t_gps) % (2.0 * np.pi)), self.latitude

    def antenna_pattern(self, right_ascension, declination, polarization, t_gps):
        """Scalar geometry helper (detector.py:268-316); the batched path computes this on the device."""
        t_gps = t_gps + self.time_delay_from_earth_center(right_ascension, declination, t_gps)
        gha = self.gmst_estimate(t_gps) - right_ascension
        cg, sg = np.cos(gha), np.sin(gha)
        cd, sd = np.cos(declination), np.sin(declination)
        cp, sp = np.cos(polarization), np.sin(polarization)
        x = np.array([-cp * sg - sp * cg * sd, -cp * cg + sp * sg * sd, sp * cd])
        y = np.array([sp * sg - cp * cg * sd, sp * cg + cp * sg * sd, cp * cd])
        dx, dy = self.response.dot(x), self.response.dot(y)
        return (x * dx - y * dy).sum(), (x * dy + y * dx).sum()

    def light_travel_time_to_detector(self, det):
        d = self.location - det.location
        return float(d.dot(d) ** 0.5 / CLIGHT_SI)

    def constants(self):
        """What the CUDA library needs from this object (bj_detector in include/bajes_b200.h)."""
        return {"response": self.response, "location": self.location, "gmst_reference": self.gmst_reference,
                "reference_time": self.reference_time, "sday": self.sday}

    # ---- measurement storage (set-up; detector.py:452-493) -------------------------------------------
    def store_measurement(self, series, noise, nspcal=0, spcal_freqs=None, nweights=0, len_weights=None):
        assert self.reference_time == series.t_gps
        self.srate = series.srate
        self.seglen = series.seglen
        self._mask = series.mask
        self._nfr = len(series.freqs)
        self.freqs = np.copy(series.freqs[series.mask])
        self.data = np.copy(series.freq_series[series.mask])
        self.psd = noise.interp_psd_pad(self.freqs) * series.window_factor
        self.nspcal, self.spcal_freqs = nspcal, spcal_freqs
        self.nweights, self.len_weights = nweights, len_weights
        self._dd = (4.0 / self.seglen) * np.sum(np.abs(self.data) ** 2.0 / self.psd)

    # ---- per-point methods: evaluated by the CUDA library --------------------------------------------
    def _need_owner(self):
        if self._owner is None:
            raise RuntimeError("this Detector is not attached to a GWLikelihood: per-point evaluations run in "
                               "the CUDA engine the likelihood owns")
        return self._owner

    def project_fdwave(self, wave, params, tag="freq", roq=None):
        """Projected strain on this detector's in-band axis (detector.py:394-418).  ``wave`` is accepted
        for signature compatibility; the engine regenerates the polarisations from ``params``."""
        if tag != "freq":
            raise NotImplementedError("time-domain approximants are outside the fused GPU path")
        return self._need_owner().project_fdwave(params)[self.ifo]

    def compute_inner_products(self, hphc, params, tag="freq", psd_weight_factor=False, roq=None):
        """(dh, hh, dd[, psd_factor]) as detector.py:496-544, with ``dh`` the SUMMED <d|h> (the reference returns the
        per-bin integrand in the non-ROQ branch because its caller sums it or Fourier-transforms it for the time
        marginalisation; here both happen on the device).  With PSD weights ``dd`` is recomputed with ``psd * weights``
        and ``psd_factor = sum(log(weights))`` (detector.py:524-531), consistent with the weighted dh / hh the engine
        returns."""
        dh, hh = self._need_owner().inner_products(params)[self.ifo]
        dd, w = self._dd, 0.0
        if self.nweights:
            weights = np.concatenate([np.full(int(n), float(params["weight{}_{}".format(b, self.ifo)]))
                                      for b, n in enumerate(self.len_weights)])          # detector.py:22-23
            w = float(np.log(weights).sum())
            dd = (4.0 / self.seglen) * float(np.sum(np.abs(self.data) ** 2.0 / (self.psd * weights)))
        if psd_weight_factor:
            return dh, hh, dd, w
        return dh, hh, dd
