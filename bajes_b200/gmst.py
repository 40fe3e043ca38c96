"""
Greenwich mean sidereal time from a GPS time stamp, dependency-free.

The reference obtains ``gmst_reference`` once per Detector from astropy
(``bajes/obs/gw/detector.py:11-15,235-242``).  astropy is not available in the
build image, so the drop-in Detector (and the import shim the oracle uses to run
the unmodified reference) falls back on this closed-form IAU-1982 expression with
UT1 ~= UTC.  It differs from astropy's IAU-2006 + IERS value by < 1e-4 rad.  That
offset is a pure sky-rotation convention: parity tests feed the *same* stored
``gmst_reference`` to the oracle and to the GPU path, so it never enters a
logL comparison.
"""
from __future__ import annotations

import math

# value of astropy.units.si.sday.si.scale (mean sidereal day in SI seconds)
SIDEREAL_DAY_S = 86164.09053083288

# GPS second at which (GPS - UTC) increments; GPS-UTC == index+1 from that second on
_LEAP_GPS = (
    46828800, 78364801, 109900802, 173059203, 252028804, 315187205, 346723206,
    393984007, 425520008, 457056009, 504489610, 551750411, 599184012, 820108813,
    914803214, 1025136015, 1119744016, 1167264017,
)

_GPS_EPOCH_JD = 2444244.5  # 1980-01-06T00:00:00 UTC


def gps_minus_utc(gps_time: float) -> int:
    n = 0
    for t in _LEAP_GPS:
        if gps_time >= t:
            n += 1
    return n


def gmst_from_gps(gps_time: float) -> float:
    """GMST in radians, in [0, 2pi)."""
    utc = float(gps_time) - gps_minus_utc(gps_time)
    days = utc / 86400.0
    d = (_GPS_EPOCH_JD - 2451545.0) + days           # days since J2000.0
    t = d / 36525.0
    # split the fast term so that no precision is lost on ~1e4 revolutions
    whole = math.floor(days)
    frac = days - whole
    d0 = (_GPS_EPOCH_JD - 2451545.0) + whole
    deg = 280.46061837 + (360.98564736629 * d0) % 360.0 \
        + 360.98564736629 * frac + 0.000387933 * t * t - t * t * t / 38710000.0
    return math.radians(deg % 360.0) % (2.0 * math.pi)
