"""
TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference's GW likelihood hot path.

This file is the ORACLE for parity tests and the ``cpu_baseline`` leg of ``bench.py``.
It must never be imported by the product package (``bajes_b200``); only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference`` legs use it.

It restates, in plain NumPy (float64 / complex128, one parameter point per call, the
same sequence of array temporaries the reference builds), the algorithm of
matteobreschi/bajes v1.1.0.  Every function cites the reference lines it follows
(paths relative to ``/root/reference``).  Parity of this restatement against the
UNMODIFIED reference (run under ``oracle/shims.py``) is pinned by
``tests/test_oracle_vs_reference.py`` (in the build container) and by the frozen vectors in
``tests/golden/`` (everywhere).  Relative binning is the one exception: the reference
module cannot be imported or run so that part is "parity unpinned".
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

# --------------------------------------------------------------------------------------
# constants  (bajes/obs/gw/approx/taylorf2.py:6-7,379 ; bajes/__init__.py:71)
# --------------------------------------------------------------------------------------
GT = 4.92549094830932e-6            # "gt": solar mass in seconds as used for v (NOT MTSUN_SI)
EULER_GAMMA = 0.57721566490153286060
PRE = 3.6686934875530996e-19        # amplitude prefactor literal
CLIGHT_SI = 299792458.0
LOG2 = 0.69314718055994528623       # literals of taylorf2.py:266-267
LOG3 = 1.0986122886681097821

Polarizations = namedtuple("Polarizations", ("plus", "cross"), defaults=([None], [None]))


# --------------------------------------------------------------------------------------
# tidal combinations  (bajes/obs/gw/utils/__init__.py:22-27,127-160)
# --------------------------------------------------------------------------------------
def lambda_tilde(m1, m2, l1, l2):
    """utils/__init__.py:127-141."""
    mt = m1 + m2
    return (16.0 / 13.0) * ((m1 + 12.0 * m2) * m1 ** 4.0 * l1 + (m2 + 12.0 * m1) * m2 ** 4.0 * l2) / mt ** 5.0


def delta_lambda_tilde(m1, m2, l1, l2):
    """utils/__init__.py:143-160 (recomputes q, eta, X from the mass fractions)."""
    mt = m1 + m2
    q = m1 / m2
    eta = q / ((1.0 + q) * (1.0 + q))
    x = np.sqrt(1.0 - 4.0 * eta)
    a4 = m1 ** 4.0
    b4 = m2 ** 4.0
    mt4 = mt ** 4.0
    first = (1690.0 * eta / 1319.0 - 4843.0 / 1319.0) * (a4 * l1 - b4 * l2) / mt4
    second = (6162.0 * x / 1319.0) * (a4 * l1 + b4 * l2) / mt4
    return first + second


def quadrupole_yy(lam):
    """Love-Q relation, utils/__init__.py:22-27."""
    if lam <= 0.0:
        return 1.0
    ll = np.log(lam)
    return np.exp(0.194 + 0.0936 * ll + 0.0474 * ll ** 2 - 4.21e-3 * ll ** 3 + 1.23e-4 * ll ** 4)


# --------------------------------------------------------------------------------------
# TaylorF2 phases and amplitude  (bajes/obs/gw/approx/taylorf2.py)
# --------------------------------------------------------------------------------------
def _vel(f, mtot):
    """v = |pi M f gt|^(1/3)  (taylorf2.py:25,178,318)."""
    return np.power(np.abs(np.pi * mtot * f * GT), 1.0 / 3.0)


def phase_tidal_6pn(f, mtot, eta, lam1, lam2):
    """taylorf2.py:9-36."""
    v = _vel(f, mtot)
    v2 = v * v
    v5 = v ** 5
    v10 = v ** 10
    v12 = v10 * v2
    dlt = np.sqrt(1.0 - 4.0 * eta)
    xa = 0.5 * (1.0 + dlt)
    xb = 0.5 * (1.0 - dlt)
    lt = lambda_tilde(xa, xb, lam1, lam2)
    dl = delta_lambda_tilde(xa, xb, lam1, lam2)
    lead = 3.0 / 128.0 / eta / v5
    return lead * (lt * v10 * (-39.0 / 2.0 - 3115.0 / 64.0 * v2) + dl * 6595.0 / 364.0 * v12)


def _tidal_75_common(f, mtot, eta, lam_a, lam_b):
    v = _vel(f, mtot)
    dlt = np.sqrt(1.0 - 4.0 * eta)
    xa = 0.5 * (1.0 + dlt)
    xb = 0.5 * (1.0 - dlt)
    pa = [xa ** n for n in range(6)]
    pb = [xb ** n for n in range(6)]
    v2 = v * v
    v3 = v2 * v
    v4 = v3 * v
    v5 = v4 * v
    kap_a = 3.0 * lam_a * pa[4] * xb
    kap_b = 3.0 * lam_b * pb[4] * xa
    n_a = -3.0 / (16.0 * eta) * (12.0 + xa / xb)
    n_b = -3.0 / (16.0 * eta) * (12.0 + xb / xa)

    def c1(p):
        return 5.0 * (3179.0 - 919.0 * p[1] - 2286.0 * p[2] + 260.0 * p[3]) / (672.0 * (12.0 - 11.0 * p[1]))

    return v2, v3, v4, v5, pa, pb, kap_a, kap_b, n_a, n_b, c1(pa), c1(pb)


def _c4_old(p):
    return -np.pi * (27719.0 - 22127.0 * p[1] + 7022.0 * p[2] - 10232.0 * p[3]) / (672.0 * (12.0 - 11.0 * p[1]))


def phase_tidal_75pn(f, mtot, eta, lam_a, lam_b):
    """taylorf2.py:38-79 (Damour-Nagar-Villain 7.5PN; all beta coefficients are zero there)."""
    v2, v3, v4, v5, pa, pb, ka, kb, na, nb, c1a, c1b = _tidal_75_common(f, mtot, eta, lam_a, lam_b)

    def c3(p):
        s = (39927845.0 / 508032.0 - 480043345.0 / 9144576.0 * p[1] + 9860575.0 / 127008.0 * p[2]
             - 421821905.0 / 2286144.0 * p[3] + 4359700.0 / 35721.0 * p[4] - 10578445.0 / 285768.0 * p[5])
        return s / (12.0 - 11.0 * p[1])

    c2 = -np.pi
    return v5 * (ka * na * (1.0 + c1a * v2 + c2 * v3 + c3(pa) * v4 + _c4_old(pa) * v5)
                 + kb * nb * (1.0 + c1b * v2 + c2 * v3 + c3(pb) * v4 + _c4_old(pb) * v5))


def phase_tidal_75pn_2020(f, mtot, eta, lam_a, lam_b):
    """taylorf2.py:82-119 (Henry-Faye-Blanchet 2020).  Quirk kept: the v^5 coefficient of
    body A uses (22415, 7598, 10520) while body B keeps (22127, 7022, 10232) (:115-116)."""
    v2, v3, v4, v5, pa, pb, ka, kb, na, nb, c1a, c1b = _tidal_75_common(f, mtot, eta, lam_a, lam_b)

    def c3(p):
        s = -5 * (-387973870.0 + 43246839.0 * p[1] + 174965616.0 * p[2] + 158378220.0 * p[3]
                  - 20427120.0 * p[4] + 4572288.0 * p[5]) / 27433728.0
        return s / (12.0 - 11.0 * p[1])

    c4a = -np.pi * (27719.0 - 22415.0 * pa[1] + 7598.0 * pa[2] - 10520.0 * pa[3]) / (672.0 * (12.0 - 11.0 * pa[1]))
    c4b = _c4_old(pb)
    c2 = -np.pi
    return v5 * (ka * na * (1.0 + c1a * v2 + c2 * v3 + c3(pa) * v4 + c4a * v5)
                 + kb * nb * (1.0 + c1b * v2 + c2 * v3 + c3(pb) * v4 + c4b * v5))


def phase_qm_35pn(f, mtot, eta, s1z, s2z, lam1, lam2):
    """taylorf2.py:121-156 (EOS-dependent self-spin term; zero when both Lambdas vanish)."""
    if lam1 == 0.0 and lam2 == 0.0:
        return 0.0
    v = _vel(f, mtot)
    v2 = v * v
    dlt = np.sqrt(1.0 - 4.0 * eta)
    x1 = 0.5 * (1.0 + dlt)
    x2 = 0.5 * (1.0 - dlt)
    a1sq = (x1 * s1z) * (x1 * s1z)
    a2sq = (x2 * s2z) * (x2 * s2z)
    q1 = quadrupole_yy(lam1) - 1.0
    q2 = quadrupole_yy(lam2) - 1.0
    plus = a1sq * q1 + a2sq * q2
    minus = a1sq * q1 - a2sq * q2
    out = -75.0 / (64.0 * eta) * plus / v
    out = out + ((45.0 / 16.0 * eta + 15635.0 / 896.0) * plus + 2215.0 / 512 * dlt * minus) * v / eta
    out = out + -75.0 / (8.0 * eta) * plus * v2 * np.pi
    return out


def phase_35pn(f, mtot, eta, s1, s2, abs_v=True):
    """Point-mass + spin 3.5PN phase, taylorf2.py:158-230 called with Lam=dLam=0 (:385-390).

    ``s1``, ``s2`` are the cartesian spin 3-vectors.
    """
    v_lso = 1.0 / np.sqrt(6.0)
    dlt = np.sqrt(1.0 - 4.0 * eta)
    v = np.abs(np.pi * mtot * f * GT) ** (1.0 / 3.0)
    v2 = v * v
    v3 = v2 * v
    v4 = v2 * v2
    v5 = v4 * v
    v6 = v3 * v3
    v7 = v3 * v4
    e2 = eta ** 2
    e3 = eta ** 3

    xa = 0.5 * (1.0 + dlt)
    xb = 0.5 * (1.0 - dlt)
    c1l, c2l = s1[2], s2[2]
    c1sq = s1[0] * s1[0] + s1[1] * s1[1] + s1[2] * s1[2]
    c2sq = s2[0] * s2[0] + s2[1] * s2[1] + s2[2] * s2[2]
    c12 = s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2]
    sl = xa * xa * c1l + xb * xb * c2l
    dsl = dlt * (xb * c2l - xa * c1l)

    sig = eta * (721.0 / 48.0 * c1l * c2l - 247.0 / 48.0 * c12)
    sig += 719.0 / 96.0 * (xa * xa * c1l * c1l + xb * xb * c2l * c2l)
    sig -= 233.0 / 96.0 * (xa * xa * c1sq + xb * xb * c2sq)
    so15 = 188.0 * sl / 3.0 + 25.0 * dsl
    gam = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * sl + (13915.0 / 84.0 - 10.0 * eta / 3.0) * dsl
    ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * c1l * c2l
    ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * xa - 120.0 * xa * xa)
            + (-4108.25 / 6.72 - 108.5 / 1.2 * xa + 125.5 / 3.6 * xa * xa)) * xa * xa * c1sq
    ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * xb - 120.0 * xb * xb)
            + (-4108.25 / 6.72 - 108.5 / 1.2 * xb + 125.5 / 3.6 * xb * xb)) * xb * xb * c2sq
    so3 = np.pi * (3760.0 * sl + 1490.0 * dsl) / 3.0 + ss3
    so35 = ((-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 - 305.0 * e2 / 36.0) * sl
            - (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 - 4735.0 * e2 / 144.0) * dsl)

    lead = 3.0 / 128.0 / eta / v5
    series = 1.0
    series += 20.0 / 9.0 * (743.0 / 336.0 + 11.0 / 4.0 * eta) * v2
    series += (so15 - 16.0 * np.pi) * v3
    series += 10.0 * (3058673.0 / 1016064.0 + 5429.0 / 1008.0 * eta + 617.0 / 144.0 * e2 - sig) * v4
    series += (38645.0 / 756.0 * np.pi - 65.0 / 9.0 * eta * np.pi - gam) * (1.0 + 3.0 * np.log(v / v_lso)) * v5
    series += (11583231236531.0 / 4694215680.0 - 640.0 / 3.0 * np.pi ** 2
               - 6848.0 / 21.0 * (EULER_GAMMA + np.log(4.0 * v))
               + (-15737765635.0 / 3048192.0 + 2255.0 * np.pi ** 2 / 12.0) * eta
               + 76055.0 / 1728.0 * e2 - 127825.0 / 1296.0 * e3 + so3) * v6
    series += (np.pi * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta - 74045.0 / 756.0 * eta ** 2) + so35) * v7
    return lead * (series + 0.0)


def _coeffs_55pn(eta):
    """Higher-order coefficients of taylorf2.py:269-295 with c_21_3PN=c_22_4PN=c_22_5PN=a_6_c=0."""
    pi2 = np.pi * np.pi
    e2 = eta * eta
    e3 = eta ** 3
    c8 = (-36946947827.5 / 1601901100.8 * eta * eta * eta * eta + 51004148102.5 / 1310646355.2 * e3
          + (30060067316599.7 / 57668439628.8 - 39954.5 / 2721.6 * pi2) * eta * eta
          + (-567987228950352.7 / 128152088064.0 - 532292.8 / 396.9 * EULER_GAMMA + 930221.5 / 5443.2 * pi2
             - 142068.8 / 44.1 * LOG2 + 2632.5 / 4.9 * LOG3) * eta
          - 9049.0 / 56.7 * pi2 - 3681.2 / 18.9 * EULER_GAMMA + 255071384399888515.3 / 83042553065472.0
          - 2632.5 / 19.6 * LOG3 - 101102.0 / 396.9 * LOG2)
    c8l = -3 * c8
    c8ll = 9 * (266146.4 / 1190.7 * eta + 1840.6 / 56.7)
    c9 = np.pi * (1032375.5 / 19958.4 * eta * eta * eta + 4529333.5 / 12700.8 * eta * eta
                  + (2255.0 / 6.0 * pi2 - 149291726073.5 / 13412044.8) * eta - 640.0 / 3.0 * pi2
                  - 1369.6 / 2.1 * EULER_GAMMA + 10534427947316.3 / 1877686272.0 - 2739.2 / 2.1 * LOG2)
    c9l = -3 * 1369.6 / 6.3 * np.pi
    c10 = 1.0 / (1.0 - 3.0 * eta) * (
        (242506658510205297979.7 / 85723270481616768.0) * eta * eta * eta * eta * eta * eta
        - (1272143474037195162.1 / 67631771583129.6) * eta * eta * eta * eta * eta
        + (1116081080066315514991.3 / 27213736660830720.0 - 943479.7 / 1881.6 * pi2) * eta * eta * eta * eta
        + (-85710407655931086054085.1 / 3428930819264670720.0 - 614779314.2 / 152806.5 * EULER_GAMMA
           - 46051.9 / 153.6 * pi2 - 4311179766.8 / 152806.5 * LOG2 + 127939.5 / 9.8 * LOG3) * eta * eta * eta
        + (-1873639936380505730110521.7 / 36575262072156487680.0 - (9923919211.9 / 458419.5) * EULER_GAMMA
           + 41579551.7 / 90316.8 * pi2 - 11734037971.3 / 458419.5 * LOG2 - 5833093.5 / 548.8 * LOG3) * eta * eta
        + (56993518125966874478111.3 / 1083711468804636672.0 + 6378740752.7 / 916839.0 * EULER_GAMMA
           - 545142954.7 / 812851.2 * pi2 + 15994339707.7 / 1833678.0 * LOG2 + (892417.5 / 313.6) * LOG3) * eta
        + (57822311.5 / 304819.2) * pi2 + (647058264.7 / 2750517.0) * EULER_GAMMA
        - 143300652329540712655.9 / 12630669799587840.0 - 551245.5 / 2195.2 * LOG3 + 5399283943.1 / 5501034.0 * LOG2)
    c10l = 3.0 / (1.0 - 3.0 * eta) * (1286378036.2 / 458419.5 * eta * eta * eta + 1384949312.9 / 1375258.5 * eta * eta
                                       - 2427943164.1 / 2750517.0 * eta + 647058264.7 / 8251551.0)
    c11 = np.pi * (65762707344.5 / 14417109907.2 * eta * eta * eta * eta - 108059782847.5 / 2621292710.4 * e3
                   + 512031495514639.7 / 62911025049.6 * e2
                   + (-1064790.5 / 3628.8 * eta * eta + 4501578.5 / 14515.2 * eta - 9439.0 / 56.7) * pi2
                   + (-134666.2 / 56.7 * EULER_GAMMA - 43038370739839704.7 / 3460106377728.0 + 2632.5 / 4.9 * LOG3
                      - 2100962.6 / 396.9 * LOG2) * eta
                   - 355801.1 / 793.8 * EULER_GAMMA + 185754140723659441.1 / 27680851021824.0
                   - 2632.5 / 19.6 * LOG3 - 86254.9 / 113.4 * LOG2)
    c11l = -3 * np.pi * (134666.2 / 170.1 * eta + 355801.1 / 2381.4)
    return c8, c8l, c8ll, c9, c9l, c10, c10l, c11, c11l


def phase_55pn(f, mtot, eta, s1, s2):
    """taylorf2.py:233-297.  Quirk kept: v is taken WITHOUT abs here (:252)."""
    v = (np.pi * mtot * f * GT) ** (1.0 / 3.0)
    v2 = v * v
    v3 = v2 * v
    v4 = v2 * v2
    v5 = v4 * v
    v7 = v3 * v4
    v8 = v7 * v
    v9 = v8 * v
    v10 = v5 * v5
    v11 = v10 * v
    lv = np.log(v)
    base = phase_35pn(f, mtot, eta, s1, s2)
    c8, c8l, c8ll, c9, c9l, c10, c10l, c11, c11l = _coeffs_55pn(eta)
    return base + (3.0 / 128.0 / eta / v5) * ((c8 + c8l * lv + c8ll * lv * lv) * v8
                                                + (c9 + c9l * lv) * v9
                                                + (c10 + c10l * lv) * v10
                                                + (c11 + c11l * lv) * v11)


def amplitude_35pn(f, mtot, eta, s1z, s2z, dist):
    """taylorf2.py:300-345."""
    mc = mtot * np.power(np.abs(eta), 3.0 / 5.0)
    dlt = np.sqrt(1.0 - 4.0 * eta)
    v = _vel(f, mtot)
    v2 = v * v
    v3 = v2 * v
    v4 = v2 * v2
    v5 = v4 * v
    v6 = v3 * v3
    v7 = v3 * v4
    e2 = eta ** 2
    e3 = eta ** 3
    a0 = (np.power(np.abs(mc), 5.0 / 6.0) / np.power(np.abs(f), 7.0 / 6.0) / dist
          / np.abs(np.pi) ** (2.0 / 3.0) * np.sqrt(5.0 / 24.0))
    cs = 0.5 * (s1z + s2z)
    ca = 0.5 * (s1z - s2z)
    be = 113.0 / 12.0 * (cs + dlt * ca - 76.0 / 113.0 * eta * cs)
    sg = ca ** 2 * (81.0 / 16.0 - 20.0 * eta) + 81.0 / 8.0 * ca * cs * dlt + cs ** 2 * (81.0 / 16.0 - eta / 4.0)
    ep = (dlt * ca * (502429.0 / 16128.0 - 907.0 / 192.0 * eta)
          + cs * (5.0 / 48.0 * e2 - 73921.0 / 2016.0 * eta + 502429.0 / 16128.0))
    return a0 * (1.0 + v2 * (11.0 / 8.0 * eta + 743.0 / 672.0) + v3 * (be / 2.0 - 2.0 * np.pi)
                 + v4 * (1379.0 / 1152.0 * e2 + 18913.0 / 16128.0 * eta + 7266251.0 / 8128512.0 - sg / 2.0)
                 + v5 * (57.0 / 16.0 * np.pi * eta - 4757.0 * np.pi / 1344.0 + ep)
                 + v6 * (856.0 / 105.0 * EULER_GAMMA + 67999.0 / 82944.0 * e3 - 1041557.0 / 258048.0 * e2
                         - 451.0 / 96.0 * np.pi ** 2 * eta + 10.0 * np.pi ** 2 / 3.0
                         + 3526813753.0 / 27869184.0 * eta - 29342493702821.0 / 500716339200.0
                         + 856.0 / 105.0 * np.log(4.0 * v))
                 + v7 * (-1349.0 / 24192.0 * e2 - 72221.0 / 24192.0 * eta - 5111593.0 / 2709504.0) * np.pi)


# approximant name -> (phase order, tidal order, new tides, quadrupole-monopole)
# (bajes/obs/gw/waveform.py:72-86 and the wrappers taylorf2.py:435-520)
APPROXIMANTS = {
    "TaylorF2_3.5PN": (7, 12, 0, 0),
    "TaylorF2_5.5PN": (11, 12, 0, 0),
    "TaylorF2_5.5PN_7.5PNTides": (11, 15, 0, 0),
    "TaylorF2_5.5PN_3.5PNQM_7.5PNTides": (11, 15, 0, 1),
    "TaylorF2_5.5PN_7.5PNTides2020": (11, 15, 1, 0),
}


def taylorf2(approx, f, p):
    """taylorf2.py:349-433 driven through the wrapper table above.  ``p`` is the parameter
    dict after ``prepare_params``.  Returns (hp, hc), complex128 arrays over ``f``."""
    order, tidal, new_tides, use_qm = APPROXIMANTS[approx]
    mtot, q = p["mtot"], p["q"]
    s1 = (p["s1x"], p["s1y"], p["s1z"])
    s2 = (p["s2x"], p["s2y"], p["s2z"])
    lam1, lam2 = p["lambda1"], p["lambda2"]
    eta = q / ((1.0 + q) * (1.0 + q))

    if order == 7:
        phi = phase_35pn(f, mtot, eta, s1, s2)
    else:
        phi = phase_55pn(f, mtot, eta, s1, s2)

    phi_t = 0.0
    if lam1 != 0.0 or lam2 != 0.0:
        if tidal == 12:
            phi_t = phase_tidal_6pn(f, mtot, eta, lam1, lam2)
        elif new_tides:
            phi_t = phase_tidal_75pn_2020(f, mtot, eta, lam1, lam2)
        else:
            phi_t = phase_tidal_75pn(f, mtot, eta, lam1, lam2)
        if use_qm:
            phi_t = phi_t + phase_qm_35pn(f, mtot, eta, s1[2], s2[2], lam1, lam2)

    phi = phi + (phi_t + p["phi_ref"] + 2 * np.pi * f * 0.0)
    amp = amplitude_35pn(f, mtot, eta, s1[2], s2[2], p["distance"])
    ci = np.cos(p["iota"])
    h = PRE * amp * np.exp(-1j * phi)
    hp = h * ((1 + ci * ci) * 0.5)
    hc = h * np.exp(-1j * np.pi / 2) * ci
    return hp, hc


# --------------------------------------------------------------------------------------
# Waveform  (bajes/obs/gw/waveform.py:253-395)
# --------------------------------------------------------------------------------------
def sph2cart(r, theta, phi):
    """bajes/pipe/__init__.py:143-149."""
    return r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)


def prepare_params(p):
    """Parameter derivations of Waveform.compute_hphc (waveform.py:334-377); mutates ``p``."""
    if "mchirp" in p:
        nu = p["q"] / (1.0 + p["q"]) ** 2.0
        p["mtot"] = p["mchirp"] / nu ** 0.6
    elif "mtot" in p:
        nu = p["q"] / (1.0 + p["q"]) ** 2.0
        p["mchirp"] = p["mtot"] * nu ** 0.6
    if "tilt1" in p and "s1" in p and "phi_1l" in p:
        p["s1x"], p["s1y"], p["s1z"] = sph2cart(p["s1"], p["tilt1"], p["phi_1l"])
    if "tilt2" in p and "s2" in p and "phi_2l" in p:
        p["s2x"], p["s2y"], p["s2z"] = sph2cart(p["s2"], p["tilt2"], p["phi_2l"])
    if "eos_logp1" in p:
        raise NotImplementedError("EOS-sampled tides (TOV solve) are outside the hot path (SURVEY.md section 2)")
    if "cos_iota" in p:
        p["iota"] = np.arccos(p["cos_iota"])
    elif "iota" in p:
        p["cos_iota"] = np.cos(p["iota"])
    return p


class WaveformPort:
    """Frequency-domain TaylorF2 generator with the interface of bajes.obs.gw.Waveform
    (waveform.py:248-395)."""

    domain = "freq"

    def __init__(self, freqs, srate, seglen, approx):
        if approx not in APPROXIMANTS:
            raise AttributeError("oracle port covers TaylorF2 approximants only: %s" % list(APPROXIMANTS))
        self.freqs = freqs
        self.srate = srate
        self.seglen = seglen
        self.approx = approx

    def compute_hphc(self, params, roq=None):
        prepare_params(params)
        with np.errstate(all="ignore"):
            hp, hc = taylorf2(self.approx, self.freqs, params)
        if not any(hp):
            return Polarizations()
        if roq is None:
            # centre the merger at seglen/2 (waveform.py:390-393)
            shift = np.exp(1j * np.pi * self.freqs * self.seglen)
            hp = hp * shift
            hc = hc * shift
        return Polarizations(plus=hp, cross=hc)


# --------------------------------------------------------------------------------------
# Noise  (bajes/obs/gw/noise.py:100-173)
# --------------------------------------------------------------------------------------
class NoisePort:
    """Padded linear interpolation of an ASD table (noise.py:100-173)."""

    def __init__(self, freqs, asd, f_min=None, f_max=None):
        freqs = np.asarray(freqs, dtype=float)
        asd = np.asarray(asd, dtype=float)
        order = np.argsort(freqs)
        self.freqs = freqs[order]
        self.df = np.median(np.diff(self.freqs))
        self.f_max = np.max(self.freqs) if f_max is None else f_max
        self.f_min = np.min(self.freqs) if f_min is None else f_min
        self.amp_spectrum = asd[order]
        self.power_spectrum = self.amp_spectrum * self.amp_spectrum
        fp, pp = self.freqs, self.power_spectrum
        if self.freqs[0] > self.df:
            n = int(round(np.min(self.freqs) / self.df)) + 1
            fp = np.append(np.flip(np.min(self.freqs) - np.arange(1, n + 1) * self.df), fp)
            pp = np.append(np.ones(n) * self.power_spectrum[0], pp)
        if self.f_max * 2 > np.max(fp):
            n = int(round((self.f_max * 2 - np.max(fp)) / self.df)) + 1
            fp = np.append(fp, np.max(fp) + np.arange(1, n + 1) * self.df)
            pp = np.append(pp, np.ones(n) * self.power_spectrum[-1])
        self.freqs_pad = fp
        self.pow_spec_pad = pp

    def interp_psd_pad(self, fr):
        """scipy interp1d(kind='linear', fill_value=inf, bounds_error=False) (noise.py:157-173)."""
        fr = np.asarray(fr, dtype=float)
        x, y = self.freqs_pad, self.pow_spec_pad
        hi = np.clip(np.searchsorted(x, fr), 1, len(x) - 1)
        lo = hi - 1
        slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
        out = slope * (fr - x[lo]) + y[lo]
        out = np.where((fr < x[0]) | (fr > x[-1]), np.inf, out)
        return out


# --------------------------------------------------------------------------------------
# Series  (bajes/obs/gw/strain.py:214-379) -- frequency-domain input only
# --------------------------------------------------------------------------------------
class SeriesPort:
    """The 'freq' branch of Series.__init__ with ``only=True`` (strain.py:330-379)."""

    def __init__(self, freq_series, freqs, srate, seglen, f_min, f_max, t_gps):
        self.window_factor = 1.0                      # strain.py:248
        self.srate = srate
        self.seglen = seglen
        self.df = 1.0 / seglen
        self.t_gps = t_gps
        self.f_min = 0.0 if f_min is None else f_min
        self.f_max = srate / 2.0 if f_max is None else f_max
        self.freqs = np.asarray(freqs)
        self.freq_series = freq_series
        assert len(self.freqs) == len(self.freq_series)
        # inclusive band mask (strain.py:375-376)
        self.mask = np.where((self.freqs >= self.f_min) & (self.freqs <= self.f_max))


# --------------------------------------------------------------------------------------
# Detector  (bajes/obs/gw/detector.py:53-544)
# --------------------------------------------------------------------------------------
# latitude, longitude, elevation, xarm_azimuth, yarm_azimuth, xarm_tilt, yarm_tilt  (detector.py:62-160)
SITES = {
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


class DetectorPort:
    """Geometry + projection + inner products of bajes.obs.gw.Detector (detector.py:164-544).

    ``gmst_reference`` / ``sday`` are taken as inputs (SURVEY.md section 8b last row): in the
    reference they come from astropy (detector.py:235-242)."""

    def __init__(self, ifo, t_gps, gmst_reference, sday):
        self.ifo = ifo
        (self.latitude, self.longitude, self.elevation, self.xarm_azimuth, self.yarm_azimuth,
         self.xarm_tilt, self.yarm_tilt) = SITES[ifo]
        self.x_arm = self._arm(self.xarm_tilt, self.xarm_azimuth)
        self.y_arm = self._arm(self.yarm_tilt, self.yarm_azimuth)
        self.location = self._location()
        # detector.py:207-211
        self.response = 0.5 * (np.einsum("i,j->ij", self.x_arm, self.x_arm)
                               - np.einsum("i,j->ij", self.y_arm, self.y_arm))
        self.reference_time = t_gps
        self.gmst_reference = gmst_reference
        self.sday = sday

    def _location(self):
        """detector.py:213-224 (WGS-84 ellipsoid)."""
        a = 6378137
        b = 6356752.314
        lat, lon, h = self.latitude, self.longitude, self.elevation
        r = a ** 2 * (a ** 2 * np.cos(lat) ** 2 + b ** 2 * np.sin(lat) ** 2) ** (-0.5)
        return np.array([(r + h) * np.cos(lat) * np.cos(lon),
                         (r + h) * np.cos(lat) * np.sin(lon),
                         ((b / a) ** 2 * r + h) * np.sin(lat)])

    def _arm(self, tilt, azimuth):
        """detector.py:226-233."""
        lat, lon = self.latitude, self.longitude
        e_lon = np.array([-np.sin(lon), np.cos(lon), 0])
        e_lat = np.array([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)])
        e_up = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        return (np.cos(tilt) * np.cos(azimuth) * e_lon + np.cos(tilt) * np.sin(azimuth) * e_lat
                + np.sin(tilt) * e_up)

    def gmst_estimate(self, gps_time):
        """detector.py:246-253."""
        dphase = (gps_time - self.reference_time) / self.sday * (2.0 * np.pi)
        return (self.gmst_reference + dphase) % (2.0 * np.pi)

    def time_delay_from_earth_center(self, ra, dec, t_gps):
        """detector.py:318-357."""
        ang = self.gmst_estimate(t_gps) - ra
        cd = np.cos(dec)
        ehat = np.array([cd * np.cos(ang), cd * -np.sin(ang), np.sin(dec)])
        dx = np.array([0, 0, 0]) - self.location
        return dx.dot(ehat) / CLIGHT_SI

    def antenna_pattern(self, ra, dec, psi, t_gps):
        """detector.py:268-316 (patterns are evaluated at t_gps + geometric delay, :286-288)."""
        t_gps = t_gps + self.time_delay_from_earth_center(ra, dec, t_gps)
        gha = self.gmst_estimate(t_gps) - ra
        cg, sg = np.cos(gha), np.sin(gha)
        cd, sd = np.cos(dec), np.sin(dec)
        cp, sp = np.cos(psi), np.sin(psi)
        x = np.array([-cp * sg - sp * cg * sd, -cp * cg + sp * sg * sd, sp * cd])
        y = np.array([sp * sg - cp * cg * sd, sp * cg + cp * sg * sd, cp * cd])
        dx = self.response.dot(x)
        dy = self.response.dot(y)
        return (x * dx - y * dy).sum(), (x * dy + y * dx).sum()

    def project_fdwave(self, wave, params, tag="freq", roq=None):
        """detector.py:394-418."""
        t = params["t_gps"] + params["time_shift"]
        fp, fc = self.antenna_pattern(params["ra"], params["dec"], params["psi"], t)
        delay = self.time_delay_from_earth_center(params["ra"], params["dec"], t) + params["time_shift"]
        proj = fp * wave.plus + fc * wave.cross
        if roq is None:
            return proj * np.exp(-2j * np.pi * self.freqs * delay)
        return proj

    def store_measurement(self, series, noise, nspcal=0, spcal_freqs=None, nweights=0, len_weights=None):
        """detector.py:452-493."""
        assert self.reference_time == series.t_gps
        self.nspcal, self.spcal_freqs = nspcal, spcal_freqs
        self.nweights, self.len_weights = nweights, len_weights
        self.srate = series.srate
        self.seglen = series.seglen
        self._mask = series.mask
        self._nfr = len(series.freqs)
        self.freqs = np.copy(series.freqs[series.mask])
        self.data = np.copy(series.freq_series[series.mask])
        self.psd = noise.interp_psd_pad(self.freqs) * series.window_factor
        self._dd = (4.0 / self.seglen) * np.sum(np.abs(self.data) ** 2.0 / self.psd)

    def compute_inner_products(self, hphc, params, tag="freq", roq=None, psd_weight_factor=False):
        """detector.py:496-544.  Returns (dh, hh, dd[, sum log w]) where ``dh`` is the per-bin integrand on the
        full axis (non-ROQ) or the scalar product (ROQ)."""
        wav = self.project_fdwave(hphc, params, tag, roq=roq)
        psd = self.psd
        dd = self._dd
        logw = 0.0
        if getattr(self, "nspcal", 0) > 0:
            # detector.py:17-20,517-519
            amp = np.array([params["spcal_amp{}_{}".format(i, self.ifo)] for i in range(self.nspcal)])
            phi = np.array([params["spcal_phi{}_{}".format(i, self.ifo)] for i in range(self.nspcal)])
            wav = wav * np.interp(self.freqs, self.spcal_freqs, (1.0 + amp) * np.exp(1j * phi))
        if getattr(self, "nweights", 0) > 0:
            # detector.py:22-23,524-531
            weights = np.concatenate([np.ones(self.len_weights[i]) * params["weight{}_{}".format(i, self.ifo)]
                                      for i in range(self.nweights)])
            logw = (np.log(weights)).sum()
            psd = psd * weights
            dd = (4.0 / self.seglen) * (np.abs(self.data) ** 2.0 / psd).sum()
        if roq is None:
            hh = (4.0 / self.seglen) * (np.abs(wav) ** 2.0 / psd).sum()
            dh = np.zeros(self._nfr, dtype=complex)
            dh[self._mask] = (4.0 / self.seglen) * np.conj(self.data) * wav / psd
            return (dh, hh, dd, logw) if psd_weight_factor else (dh, hh, dd)
        if roq is not None:
            hh = np.dot(roq[self.ifo]["psi_weights"], np.abs(wav[roq["mask_psi"]]) ** 2)
            tc = (self.seglen / 2.0
                  + self.time_delay_from_earth_center(params["ra"], params["dec"],
                                                      params["t_gps"] + params["time_shift"])
                  + params["time_shift"])
            dh = np.dot(roq[self.ifo]["omega_weights_interp"](tc), wav[roq["mask_omega"]])
        else:
            hh = (4.0 / self.seglen) * (np.abs(wav) ** 2.0 / self.psd).sum()
            dh = np.zeros(self._nfr, dtype=complex)
            dh[self._mask] = (4.0 / self.seglen) * np.conj(self.data) * wav / self.psd
        return dh, hh, self._dd


# --------------------------------------------------------------------------------------
# GWLikelihood  (bajes/pipe/log_like.py:21-183)
# --------------------------------------------------------------------------------------
def _log_i0(x):
    from scipy.special import i0e
    return np.log(i0e(x)) + x


class GWLikelihoodPort:
    """Restatement of GWLikelihood (pipe/log_like.py:21-183) without spcal / PSD weights."""

    def __init__(self, ifos, datas, dets, noises, freqs, srate, seglen, approx,
                 marg_phi_ref=False, marg_time_shift=False, roq=None,
                 nspcal=0, spcal_freqs=None, nweights=0, len_weights=None):
        self.ifos = ifos
        self.dets = dets
        self.roq = roq
        self.marg_phi_ref = marg_phi_ref
        self.marg_time_shift = marg_time_shift
        if roq is not None and marg_time_shift:
            raise AttributeError("Time-shift marginalization has not been implemented with the ROQ approximation.")
        self.logZ_noise = 0.0
        mask = None
        for ifo in ifos:
            self.dets[ifo].store_measurement(datas[ifo], noises[ifo], nspcal, spcal_freqs, nweights, len_weights)
            if datas[ifo].seglen != seglen:
                raise ValueError("Input seglen of data and model do not match in detector {}.".format(ifo))
            self.logZ_noise += -0.5 * self.dets[ifo]._dd
            self.Nfr = len(datas[ifo].freqs)
            mask = datas[ifo].mask
        if roq is not None:
            self.wave = WaveformPort(roq["freqs_join"], srate, seglen, approx)
        else:
            self.wave = WaveformPort(freqs[mask], srate, seglen, approx)

    def inner_products(self, params):
        """Per-detector (sum dh, hh); helper for tests, not in the reference."""
        wave = self.wave.compute_hphc(params, roq=self.roq)
        out = {}
        for ifo in self.ifos:
            dh, hh, _ = self.dets[ifo].compute_inner_products(wave, params, "freq", roq=self.roq)
            out[ifo] = (dh if self.roq is not None else dh.sum(), float(np.real(hh)))
        return out

    def log_like(self, params):
        from scipy.special import logsumexp
        wave = self.wave.compute_hphc(params, roq=self.roq)
        # log_like.py:121-129
        if not any(wave.plus):
            return -np.inf
        if np.any(np.isnan(wave.plus)) or np.any(np.isnan(wave.cross)):
            return -np.inf
        if np.any(np.isinf(wave.plus)) or np.any(np.isinf(wave.cross)):
            return -np.inf
        hh = 0.0
        dd = 0.0
        psd_fact = 0.0
        if self.marg_time_shift:
            # log_like.py:135-156
            arr = np.zeros(self.Nfr, dtype=complex)
            for ifo in self.ifos:
                dh_i, hh_i, dd_i, lw = self.dets[ifo].compute_inner_products(wave, params, "freq",
                                                                               psd_weight_factor=True)
                arr = arr + np.fft.fft(dh_i)
                hh += np.real(hh_i)
                dd += np.real(dd_i)
                psd_fact += lw
            if self.marg_phi_ref:
                r = logsumexp(_log_i0(np.abs(arr)) - np.log(self.Nfr))
            else:
                r = logsumexp(np.real(arr) - np.log(self.Nfr))
        else:
            # log_like.py:158-179
            dh = 0.0 + 0.0j
            for ifo in self.ifos:
                if self.roq is not None:
                    dh_i, hh_i, dd_i = self.dets[ifo].compute_inner_products(wave, params, "freq", roq=self.roq)
                    lw = 0.0
                else:
                    dh_i, hh_i, dd_i, lw = self.dets[ifo].compute_inner_products(wave, params, "freq",
                                                                                   psd_weight_factor=True)
                dh += dh_i if self.roq is not None else dh_i.sum()
                hh += np.real(hh_i)
                dd += np.real(dd_i)
                psd_fact += lw
            r = _log_i0(np.abs(dh)) if self.marg_phi_ref else np.real(dh)
        return -0.5 * (hh + dd) + r - self.logZ_noise - 0.5 * psd_fact


# --------------------------------------------------------------------------------------
# ROQ set-up  (bajes/pipe/utils/roq.py:808-872, bajes/pipe/__init__.py:661-716,753-783)
# --------------------------------------------------------------------------------------
def roq_linear_weights(tc, df, basis_linear, data, psd, freqs):
    """roq.py:808-815  (eq. 9b of arXiv:1604.08253)."""
    vec = np.conj(data) / psd * np.exp(-2.0 * np.pi * 1j * tc * freqs)
    return 4.0 * df * (basis_linear.T @ vec)


def roq_setup_one_ifo(lin_freqs, quad_freqs, basis_linear, basis_quadratic,
                      freqs_full, data_f, psd, seglen, t_low, t_high, time_points):
    """roq.py:817-872 + pipe/__init__.py:689,706-714 for one detector, with the basis passed in
    memory (JenpyROQ layout: interpolants are [N_freq, N_nodes])."""
    from scipy.interpolate import interp1d
    tcs = np.linspace(seglen / 2.0 + t_low - 0.030, seglen / 2.0 + t_high + 0.030, time_points)
    df = 1.0 / seglen
    rows = [roq_linear_weights(tc, df, basis_linear, data_f, psd, freqs_full) for tc in tcs]
    matrix = np.transpose(rows)
    omega = interp1d(tcs, matrix, kind="cubic", bounds_error=False, fill_value=-np.inf)
    psi = 4.0 * df * basis_quadratic.T @ (1.0 / psd)
    joined = np.sort(np.array(list(set(np.concatenate((quad_freqs, lin_freqs))))))
    mask_lin = [i for i in range(len(joined)) if joined[i] in lin_freqs]
    mask_qua = [i for i in range(len(joined)) if joined[i] in quad_freqs]
    return joined, mask_qua, mask_lin, psi, omega


# --------------------------------------------------------------------------------------
# Relative binning  (bajes/pipe/utils/binning.py) -- PARITY UNPINNED (reference is dead code, F5)
# --------------------------------------------------------------------------------------
def relbin_setup_bins(f_full, f_lo, f_hi, chi=1.0, eps=0.5):
    """binning.py:22-49."""
    f = np.linspace(f_lo, f_hi, 10000)
    ga = np.array([-5.0 / 3.0, -2.0 / 3.0, 1.0, 5.0 / 3.0, 7.0 / 3.0])
    dalp = chi * 2.0 * np.pi / np.absolute(f_lo ** ga - f_hi ** ga)
    dphi = np.sum(np.array([np.sign(ga[i]) * dalp[i] * f ** ga[i] for i in range(len(ga))]), axis=0)
    big = dphi - dphi[0]
    nbin = int(big[-1] // eps)
    grid = np.linspace(big[0], big[-1], nbin + 1)
    fbin = np.interp(grid, big, f)            # 'slinear' interp1d on a monotone abscissa
    idx = np.array([np.argmin(np.absolute(f_full - ff)) for ff in fbin])
    fbin = np.array([f_full[i] for i in idx])
    return nbin, fbin, idx


def relbin_summary_data(f, fbin, idx, psds, datas, h0s):
    """binning.py:52-101.  Returns array [4, n_det, n_bin]: A0, A1, B0, B1."""
    nbin = len(fbin) - 1
    T = 1.0 / np.median(np.diff(f))
    out = np.zeros((4, len(psds), nbin), dtype=complex)
    for k in range(len(psds)):
        for i in range(nbin):
            sl = slice(idx[i], idx[i + 1])
            fc = f[sl] - 0.5 * (fbin[i] + fbin[i + 1])
            dh0 = datas[k][sl] * np.conjugate(h0s[k][sl]) / psds[k][sl]
            hh0 = np.absolute(h0s[k][sl]) ** 2 / psds[k][sl]
            out[0, k, i] = (4.0 / T) * np.sum(dh0)
            out[1, k, i] = (4.0 / T) * np.sum(dh0 * fc)
            out[2, k, i] = (4.0 / T) * np.sum(hh0)
            out[3, k, i] = (4.0 / T) * np.sum(hh0 * fc)
    return out


class GWBinningLikelihoodPort:
    """Restatement of GWBinningLikelihood (binning.py:105-306) with the documented repairs
    (SURVEY.md section 8c): polarisations passed as a tuple type, fiducial waveform built on the
    SAME masked axis as the detector data, and the per-bin contributions are SUMMED (the
    Zackay-Dai-Venumadhav estimator) instead of ``trapz(.., x=fax)`` (:284)."""

    def __init__(self, ifos, datas, dets, noises, fiducial_params, freqs, srate, seglen, approx,
                 marg_phi_ref=False, chi=1.0, eps=0.5):
        self.ifos = ifos
        self.dets = dets
        self.seglen = seglen
        self.marg_phi_ref = marg_phi_ref
        mask = datas[ifos[0]].mask
        fband = np.asarray(freqs)[mask]
        wave0 = WaveformPort(fband, srate, seglen, approx)
        h0 = wave0.compute_hphc(dict(fiducial_params))
        psds, sfts, h0s = [], [], []
        for ifo in ifos:
            dets[ifo].store_measurement(datas[ifo], noises[ifo])
            psds.append(dets[ifo].psd)
            sfts.append(dets[ifo].data)
            h0s.append(dets[ifo].project_fdwave(h0, fiducial_params, "freq"))
        f_min, f_max = datas[ifos[0]].f_min, datas[ifos[0]].f_max
        self.Nbin, self.fbin, self.fbin_ind = relbin_setup_bins(fband, f_min, f_max, chi=chi, eps=eps)
        self.sdat = relbin_summary_data(fband, self.fbin, self.fbin_ind, psds, sfts, h0s)
        self.h0_bin = np.array([h[self.fbin_ind] for h in h0s])
        self.wave = WaveformPort(self.fbin, srate, seglen, approx)
        for ifo in ifos:
            dets[ifo].freqs = self.fbin

    def inner_products(self, params):
        hphc = self.wave.compute_hphc(params)
        dh_tot, hh_tot = 0.0 + 0.0j, 0.0
        for k, ifo in enumerate(self.ifos):
            h = self.dets[ifo].project_fdwave(hphc, params, "freq")
            r = h / self.h0_bin[k]                                # binning.py:301-304
            r0 = 0.5 * (r[:-1] + r[1:])
            r1 = (r[1:] - r[:-1]) / (self.fbin[1:] - self.fbin[:-1])
            hh = self.sdat[2][k] * np.absolute(r0) ** 2 + self.sdat[3][k] * 2.0 * (r0 * np.conjugate(r1)).real
            dh = self.sdat[0][k] * np.conjugate(r0) + self.sdat[1][k] * np.conjugate(r1)
            dh_tot += np.sum(dh)
            hh_tot += np.real(np.sum(hh))
        return dh_tot, hh_tot

    def log_like(self, params):
        """binning.py:248-271."""
        dh, hh = self.inner_products(params)
        if not (np.isfinite(dh.real) and np.isfinite(dh.imag) and np.isfinite(hh)):
            return -np.inf
        r = _log_i0(np.abs(dh)) if self.marg_phi_ref else np.real(dh)
        return float(np.real(r - 0.5 * hh))


# ---------------------------------------------------------------------------------------------------------
# dynesty random-walk proposal (bajes/inf/sampler/dynesty.py:503-614), ONE chain, sequential -- the checker
# of the product's lock-step version (bajes_b200/rwalk.py).  ``rstate`` stands for the global ``numpy.random``
# the reference draws from (same legacy generator, same call sequence).
# ---------------------------------------------------------------------------------------------------------

def _nmcmc_port(accept_ratio, old_act, maxmcmc, safety=5):
    """inf/utils.py:140-161."""
    tau = maxmcmc / safety
    if accept_ratio == 0.0:
        exact = (1 + 1 / tau) * old_act
    else:
        exact = float(min((1.0 - 1.0 / tau) * old_act + (safety / tau) * (2.0 / accept_ratio - 1.0), maxmcmc))
    return max(safety, int(exact))


def _inside_port(u, nonbounded):
    """dynesty.py:27-42."""
    if nonbounded is not None and not np.any(nonbounded):
        nonbounded = None
    if nonbounded is None:
        return u.min() > 0 and u.max() < 1
    hard, soft = u[nonbounded], u[~nonbounded]
    return hard.min() > 0 and hard.max() < 1 and soft.min() > -0.5 and soft.max() < 1.5


def sample_rwalk_port(args, maxmcmc, walks, nact, rstate):
    u, loglstar, axes, scale, prior_transform, loglikelihood, _seedseq, kwargs = args
    nonbounded, periodic, reflective = kwargs.get("nonbounded"), kwargs.get("periodic"), kwargs.get("reflective")
    n = len(u)
    n_acc = n_rej = n_out = 0
    act = np.inf
    chain = []                                   # (u, v, logl) per recorded step
    here = None
    while 0.5 * nact * act >= len(chain):
        direction = rstate.randn(n)
        direction /= np.linalg.norm(direction)
        step = direction * rstate.rand() ** (1.0 / n)
        trial = u + scale * np.dot(axes, step)
        if periodic is not None:
            trial[periodic] = np.mod(trial[periodic], 1)
        if reflective is not None:
            x = trial[reflective]
            even = np.mod(x, 2) < 1
            trial[reflective] = np.where(even, np.mod(x, 1), 1 - np.mod(x, 1))
        if not _inside_port(trial, nonbounded):
            n_out += 1
            if n_acc > 0:
                chain.append(here)
            continue
        v_trial = prior_transform(np.array(trial))
        l_trial = loglikelihood(np.array(v_trial))
        if l_trial > loglstar:
            u = trial
            here = (trial, v_trial, l_trial)
            n_acc += 1
            chain.append(here)
        else:
            n_rej += 1
            if n_acc > 0:
                chain.append(here)
        if n_acc + n_rej > walks and n_acc > 1:
            act = _nmcmc_port(n_acc / (n_acc + n_rej + n_out), walks, maxmcmc)
        if n_acc + n_rej > maxmcmc:
            break
    if np.isfinite(act) and int(0.5 * nact * act) < len(chain):
        u_out, v_out, l_out = chain[rstate.randint(int(0.5 * nact * act), len(chain))]
    else:
        u_out = rstate.uniform(size=n)
        v_out = prior_transform(u_out)
        l_out = loglikelihood(v_out)
    return u_out, v_out, l_out, n_acc + n_rej, {"accept": n_acc, "reject": n_rej, "fail": n_out, "scale": scale}
