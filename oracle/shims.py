"""
TEST INFRASTRUCTURE -- not part of the product path.

Process-local compatibility shims that let the UNMODIFIED reference package
(``/root/reference/bajes``) import under NumPy 2.3 / SciPy 1.18 with no astropy /
matplotlib installed.  Nothing under ``/root/reference`` is
edited or copied: we only register aliases / stub modules in ``sys.modules`` before
``import bajes``.

Used only by ``oracle/make_golden.py`` and by the ``needs_reference`` tests (which are
skipped when ``/root/reference`` does not exist, e.g. on the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BAJES_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bajes"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__["__path__"] = []
    sys.modules[name] = m
    return m


def install() -> None:
    """Install the shims (idempotent) and put the reference on sys.path."""
    import numpy as np
    import scipy.integrate
    import scipy.signal
    import scipy.signal.windows

    # bajes/obs/gw/strain.py:3  ``from scipy.signal import tukey``
    if not hasattr(scipy.signal, "tukey"):
        scipy.signal.tukey = scipy.signal.windows.tukey
    # bajes/pipe/utils/binning.py:13  ``from scipy.integrate import trapz``
    if not hasattr(scipy.integrate, "trapz"):
        scipy.integrate.trapz = scipy.integrate.trapezoid
    # removed NumPy aliases (obs/gw/utils/__init__.py:315,321, inf/prior.py:290, ptmcmc.py:259-278)
    for alias, target in (("int", int), ("float", float), ("complex", complex), ("bool", bool)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    if "trapz" not in np.__dict__:
        np.trapz = np.trapezoid
    # inf/sampler/proposal.py:116 (and the other proposals): ``q, fact = np.transpose(q_f)`` on a ragged list of
    # (array, float) pairs -- an implicit object array before NumPy 1.24, a ValueError since
    if not getattr(np.transpose, "_bajes_shim", False):
        _transpose = np.transpose

        def transpose(a, axes=None):
            try:
                return _transpose(a, axes)
            except ValueError:
                rows = list(a)
                out = np.empty((len(rows[0]), len(rows)), dtype=object)
                for i, r in enumerate(rows):
                    for j, v in enumerate(r):
                        out[j, i] = v
                return out

        transpose._bajes_shim = True
        np.transpose = transpose

    # bajes/obs/utils/cosmo.py:9  ``from scipy.misc import derivative``
    try:
        import scipy.misc as _misc  # noqa: F401
        if not hasattr(_misc, "derivative"):
            raise ImportError
    except Exception:
        misc = sys.modules.get("scipy.misc") or _stub("scipy.misc")

        def derivative(func, x0, dx=1.0, n=1, args=(), order=3):
            return (func(x0 + dx, *args) - func(x0 - dx, *args)) / (2.0 * dx)

        misc.derivative = derivative
        import scipy
        scipy.misc = misc

    # matplotlib (obs/utils/cosmo.py:4, obs/gw/noise.py:217)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib")
            mpl.use = lambda *a, **k: None
            plt = _stub("matplotlib.pyplot")
            mpl.pyplot = plt

    # astropy (obs/gw/detector.py:11-15,240-241; obs/utils/cosmo.py:3)
    if "astropy" not in sys.modules:
        try:
            import astropy  # noqa: F401
        except Exception:
            _install_astropy_stub()

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def _install_astropy_stub() -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from bajes_b200.gmst import SIDEREAL_DAY_S, gmst_from_gps

    astropy = _stub("astropy")
    units = _stub("astropy.units")
    si = _stub("astropy.units.si")
    time = _stub("astropy.time")
    cosmology = _stub("astropy.cosmology")

    class _Scale:
        def __init__(self, scale):
            self.scale = scale

    class _Unit:
        def __init__(self, scale):
            self.si = _Scale(scale)

        def __rmul__(self, other):
            return other

        def __mul__(self, other):
            return other

    si.sday = _Unit(SIDEREAL_DAY_S)
    units.si = si
    units.sday = si.sday
    for name in ("Mpc", "km", "s", "m", "K", "eV", "Gyr"):
        setattr(units, name, _Unit(1.0))

    class _Angle:
        def __init__(self, rad):
            self.rad = rad

    class Time:
        def __init__(self, val, format=None, scale=None, location=None):
            self.val = val

        def sidereal_time(self, kind, longitude=None):
            return _Angle(gmst_from_gps(self.val))

    time.Time = Time

    class _Cosmo:
        def __init__(self, *a, **k):
            pass

    for name in ("FlatLambdaCDM", "LambdaCDM", "Planck15", "Planck18", "z_at_value"):
        setattr(cosmology, name, _Cosmo)

    astropy.units = units
    astropy.time = time
    astropy.cosmology = cosmology


def import_reference():
    """Return the imported (unmodified) reference ``bajes`` package."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import bajes  # noqa: F401
    return bajes
