"""
ctypes binding of ``libbajes_b200.so`` (C ABI in ``include/bajes_b200.h``).

This is the binding a bajes maintainer would add: bajes is pure Python, so
its "FFI" for the likelihood path is ctypes over plain pointers.  There is NO CPU fallback:
a missing library or a missing B200 raises ``EngineUnavailable`` loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

BJ_ABI_VERSION = 2
BJ_MAX_IFO = 3
BJ_MAX_SPCAL = 16
BJ_MAX_WEIGHTS = 16
BJ_COL_CONST = -1
BJ_COL_ABSENT = -2

PARAM_SLOTS = (
    "mchirp", "mtot", "q",
    "s1x", "s1y", "s1z", "s2x", "s2y", "s2z",
    "s1", "tilt1", "phi_1l", "s2", "tilt2", "phi_2l",
    "lambda1", "lambda2",
    "distance", "cos_iota", "iota",
    "ra", "dec", "psi", "time_shift", "phi_ref", "t_gps",
)
BJ_NUM_PARAMS = len(PARAM_SLOTS)
SLOT_INDEX = {n: i for i, n in enumerate(PARAM_SLOTS)}

# bajes/obs/gw/waveform.py:72-86
APPROX_IDS = {
    "TaylorF2_3.5PN": 0,
    "TaylorF2_5.5PN": 1,
    "TaylorF2_5.5PN_7.5PNTides": 2,
    "TaylorF2_5.5PN_3.5PNQM_7.5PNTides": 3,
    "TaylorF2_5.5PN_7.5PNTides2020": 4,
}
SINCOS_MODES = {"mufu": 0, "poly32": 1, "fp64": 2}

LIB_NAME = "libbajes_b200.so"
LIB_PATH = os.environ.get("BAJES_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)


class EngineUnavailable(RuntimeError):
    """The CUDA library or a usable B200 is missing.  There is deliberately no CPU fallback."""


class EngineError(RuntimeError):
    pass


class bj_param_map(C.Structure):
    _fields_ = [("column", C.c_int32 * BJ_NUM_PARAMS), ("value", C.c_double * BJ_NUM_PARAMS)]


class bj_detector(C.Structure):
    _fields_ = [("response", C.c_double * 9), ("location", C.c_double * 3),
                ("gmst_reference", C.c_double), ("reference_time", C.c_double), ("sday", C.c_double)]


class bj_config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("approx", C.c_int32),
                ("n_ifo", C.c_int32), ("marg_phi_ref", C.c_int32), ("sincos", C.c_int32),
                ("seglen", C.c_double), ("dd_sum", C.c_double), ("logZ_noise", C.c_double),
                ("det", bj_detector * BJ_MAX_IFO)]


class bj_extras(C.Structure):
    _fields_ = [("nspcal", C.c_int32), ("nweights", C.c_int32),
                ("spcal_freqs", C.POINTER(C.c_double)), ("len_weights", C.POINTER(C.c_int64)),
                ("marg_time_shift", C.c_int32), ("reserved", C.c_int32),
                ("n_full", C.c_int64), ("band_offset", C.c_int64),
                ("compress", C.c_int32), ("reserved2", C.c_int32),
                ("compress_mchirp_min", C.c_double), ("compress_tau_max", C.c_double)]


class bj_extra_map(C.Structure):
    _fields_ = [("spcal_amp_col", (C.c_int32 * BJ_MAX_SPCAL) * BJ_MAX_IFO),
                ("spcal_phi_col", (C.c_int32 * BJ_MAX_SPCAL) * BJ_MAX_IFO),
                ("weight_col", (C.c_int32 * BJ_MAX_WEIGHTS) * BJ_MAX_IFO),
                ("spcal_amp_val", (C.c_double * BJ_MAX_SPCAL) * BJ_MAX_IFO),
                ("spcal_phi_val", (C.c_double * BJ_MAX_SPCAL) * BJ_MAX_IFO),
                ("weight_val", (C.c_double * BJ_MAX_WEIGHTS) * BJ_MAX_IFO)]


class bj_roq_tables(C.Structure):
    _fields_ = [("n_join", C.c_int64), ("freqs_join", C.POINTER(C.c_double)),
                ("n_lin", C.c_int64), ("mask_omega", C.POINTER(C.c_int32)),
                ("n_quad", C.c_int64), ("mask_psi", C.POINTER(C.c_int32)),
                ("n_knots", C.c_int64), ("n_coef", C.c_int64),
                ("knots", C.POINTER(C.POINTER(C.c_double))),
                ("coef_ri", C.POINTER(C.POINTER(C.c_double))),
                ("psi_weights", C.POINTER(C.POINTER(C.c_double))),
                ("tc_min", C.c_double), ("tc_max", C.c_double)]


_DP = C.POINTER(C.c_double)
_DPP = C.POINTER(_DP)

# every symbol include/bajes_b200.h declares: name -> (restype, argtypes)
EXPORTS = {
    "bj_last_error": (C.c_char_p, []),
    "bj_abi_version": (C.c_int, []),
    "bj_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bj_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(bj_config), C.c_int64, _DP, _DPP, _DPP]),
    "bj_create_ex": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(bj_config), C.c_int64, _DP, _DPP, _DPP,
                               C.POINTER(bj_extras)]),
    "bj_destroy": (C.c_int, [C.c_void_p]),
    "bj_loglike_ex": (C.c_int, [C.c_void_p, _DP, C.c_int64, C.c_int32, C.POINTER(bj_param_map),
                                C.POINTER(bj_extra_map), _DP]),
    "bj_loglike_device_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(bj_param_map),
                                       C.POINTER(bj_extra_map), C.c_void_p, C.c_void_p, C.c_void_p]),
    "bj_loglike": (C.c_int, [C.c_void_p, _DP, C.c_int64, C.c_int32, C.POINTER(bj_param_map), _DP]),
    "bj_loglike_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(bj_param_map),
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "bj_num_chunks": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "bj_partial_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(bj_param_map),
                                    C.POINTER(bj_extra_map), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bj_finish_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bj_project_waveform": (C.c_int, [C.c_void_p, _DP, C.c_int32, C.POINTER(bj_param_map), _DP]),
    "bj_polarizations": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int64, _DP, _DP, C.c_int32,
                                   C.POINTER(bj_param_map), _DP, _DP]),
    "bj_roq_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(bj_config), C.POINTER(bj_roq_tables)]),
    "bj_relbin_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(bj_config), C.c_int64, _DP, _DPP, _DP]),
    "bj_peer_create": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "bj_peer_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "bj_peer_close": (C.c_int, [C.c_int, C.c_void_p, C.c_int]),
    "bj_peer_signal": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "bj_peer_scatter": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_int64, C.c_void_p]),
    "bj_peer_wait": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_double, C.c_void_p]),
    "bj_wavebank": (C.c_int, [C.c_int, C.c_int, C.c_int64, _DP, _DP, C.c_int64, C.c_int32, C.POINTER(bj_param_map), C.c_int32,
                              C.c_void_p]),
    "bj_rb_greedy": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, _DP, C.c_double, C.c_double, C.c_int32, C.c_int32,
                               C.c_void_p, C.POINTER(C.c_int32), _DP, C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "bj_measure_peak": (C.c_int, [C.c_int, C.c_int, _DP]),
    "bj_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "bj_last_launch_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "bj_last_tile_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "bj_table_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "bj_compress_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), _DP]),
    "bj_last_level_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "bj_last_levels": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    "bj_focus": (C.c_int, [C.c_void_p, _DP, C.c_int32, C.POINTER(bj_param_map), C.c_double, C.c_double, C.c_double,
                           C.POINTER(C.c_int32)]),
    "bj_unfocus": (C.c_int, [C.c_void_p]),
    "bj_compress_axis_host": (C.c_int, [C.c_int64, _DP, C.c_int32, _DPP, C.c_int32, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.POINTER(C.c_int64), _DP, _DPP, C.POINTER(C.c_int32)]),
}

_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the in-tree library and bind every exported symbol (raises if any is missing)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineUnavailable(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  bajes_b200 has no CPU fallback." % p)
    try:
        lib = C.CDLL(p)
    except OSError as exc:  # pragma: no cover
        raise EngineUnavailable("cannot load %s: %s" % (p, exc)) from exc
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.bj_abi_version() != BJ_ABI_VERSION:
        raise EngineUnavailable("ABI mismatch: library %d, binding %d" % (lib.bj_abi_version(), BJ_ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc == 0:
        return
    msg = lib.bj_last_error().decode("utf-8", "replace")
    if rc == -4:
        raise EngineUnavailable("%s: %s" % (what, msg))
    if rc == -3:
        raise NotImplementedError("%s: %s" % (what, msg))
    raise EngineError("%s failed (status %d): %s" % (what, rc, msg))


def device_info(device: int = 0) -> Dict[str, object]:
    lib = load_library()
    name = C.create_string_buffer(128)
    sm, maj, mnr, khz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib, lib.bj_device_info(device, name, 128, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(khz)),
           "bj_device_info")
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value), "clock_khz": khz.value}


def measure_peak(which: str, device: int = 0) -> float:
    """Issue-rate micro-benchmark: 'dfma' or 'mufu' thread-instructions per second."""
    lib = load_library()
    out = C.c_double()
    _check(lib, lib.bj_measure_peak(device, {"dfma": 0, "mufu": 1}[which], C.byref(out)), "bj_measure_peak")
    return out.value


PEER_FLAGS = 64
PEER_HEADER_BYTES = (PEER_FLAGS + 2) * 8


class PeerBuffer:
    """The sampler rank's result buffer, written by every rank's epilogue kernel over NVLink
    (include/bajes_b200.h: bj_peer_*).  ``create`` on the owner, ``open`` with its 64-byte handle elsewhere."""

    def __init__(self, device: int, base: int, opened: bool, capacity: int, handle: bytes = b""):
        self.device, self.base, self.opened, self.capacity, self.handle = device, base, opened, capacity, handle

    @classmethod
    def create(cls, device: int, capacity: int) -> "PeerBuffer":
        lib = load_library()
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        _check(lib, lib.bj_peer_create(device, capacity, C.byref(base), handle), "bj_peer_create")
        return cls(device, base.value, False, capacity, handle.raw)

    @classmethod
    def open(cls, device: int, handle: bytes, capacity: int) -> "PeerBuffer":
        lib = load_library()
        base = C.c_void_p()
        _check(lib, lib.bj_peer_open(device, handle, C.byref(base)), "bj_peer_open")
        return cls(device, base.value, True, capacity, handle)

    @property
    def data_ptr(self) -> int:
        """device address of logL[0]"""
        return self.base + PEER_HEADER_BYTES

    def signal(self, slot: int, value: int, stream: int = 0) -> None:
        lib = load_library()
        _check(lib, lib.bj_peer_signal(self.device, C.c_void_p(self.base), slot, value, C.c_void_p(stream or None)),
               "bj_peer_signal")

    def scatter(self, dst_offset_bytes: int, src_ptr: int, stage_ptr: int, begins, ends, slots, value: int,
                stream: int = 0) -> None:
        """byte ranges [begins[p], ends[p]) of the host block at src_ptr -> this buffer (+ dst_offset_bytes), each
        followed by its flag; stage_ptr = 0 when the source is pinned (bj_peer_scatter)"""
        lib = load_library()
        b = np.ascontiguousarray(begins, dtype=np.int64)
        e = np.ascontiguousarray(ends, dtype=np.int64)
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        _check(lib, lib.bj_peer_scatter(self.device, C.c_void_p(self.base), dst_offset_bytes, C.c_void_p(src_ptr),
                                        C.c_void_p(stage_ptr or None), len(b), b.ctypes.data_as(C.POINTER(C.c_int64)),
                                        e.ctypes.data_as(C.POINTER(C.c_int64)), sl.ctypes.data_as(C.POINTER(C.c_int32)),
                                        value, C.c_void_p(stream or None)), "bj_peer_scatter")

    def wait(self, first: int, n: int, value: int, timeout_ms: float = 5000.0, stream: int = 0) -> None:
        lib = load_library()
        _check(lib, lib.bj_peer_wait(self.device, C.c_void_p(self.base), first, n, value, float(timeout_ms),
                                     C.c_void_p(stream or None)), "bj_peer_wait")

    def close(self) -> None:
        if self.base:
            load_library().bj_peer_close(self.device, C.c_void_p(self.base), 1 if self.opened else 0)
            self.base = 0


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def _ptr_array(arrs: Sequence[np.ndarray]):
    out = (_DP * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = _ptr(a)
    return out


class ParamMap:
    """names -> columns of the batch array, plus constants (replaces Prior.this_sample +
    the dict parsing in Waveform.compute_hphc; bajes/inf/prior.py:262-269, waveform.py:334-377)."""

    UNSUPPORTED_PREFIXES = ("eos_",)
    EXTRA_PREFIXES = ("spcal_amp", "spcal_phi", "weight")

    def __init__(self, names: Sequence[str], const: Optional[Dict[str, float]] = None, ifos: Sequence[str] = (),
                 nspcal: int = 0, nweights: int = 0):
        """``ifos`` / ``nspcal`` / ``nweights``: when the likelihood was built with calibration envelopes or PSD
        weights, the parameters 'spcal_amp{j}_{ifo}', 'spcal_phi{j}_{ifo}', 'weight{b}_{ifo}'
        (bajes/obs/gw/detector.py:17-23) are looked up among the sampled names, then the constants."""
        const = dict(const or {})
        self.names = list(names)
        self.ndim = len(self.names)
        for n in list(self.names) + list(const):
            if n.startswith(self.UNSUPPORTED_PREFIXES):
                raise NotImplementedError(
                    "parameter '%s' (EOS sampling) is outside the fused GPU likelihood; there is no CPU fallback" % n)
            if n.startswith(self.EXTRA_PREFIXES) and n != "weights" and not (nspcal or nweights):
                raise NotImplementedError(
                    "parameter '%s' given but the likelihood was built without calibration envelopes / PSD weights "
                    "(nspcal = nweights = 0)" % n)
        self.extra = None
        if nspcal or nweights:
            self.extra = bj_extra_map()
            col = {n: i for i, n in enumerate(self.names)}

            def place(cols, vals, i, j, key):
                if key in col:
                    cols[i][j] = col[key]
                elif key in const:
                    cols[i][j] = BJ_COL_CONST
                    vals[i][j] = float(const[key])
                else:
                    raise KeyError("parameter '%s' is neither sampled nor constant" % key)

            for i, ifo in enumerate(ifos):
                for j in range(nspcal):
                    place(self.extra.spcal_amp_col, self.extra.spcal_amp_val, i, j, "spcal_amp{}_{}".format(j, ifo))
                    place(self.extra.spcal_phi_col, self.extra.spcal_phi_val, i, j, "spcal_phi{}_{}".format(j, ifo))
                for b in range(nweights):
                    place(self.extra.weight_col, self.extra.weight_val, i, b, "weight{}_{}".format(b, ifo))
        self.c = bj_param_map()
        for i in range(BJ_NUM_PARAMS):
            self.c.column[i] = BJ_COL_ABSENT
            self.c.value[i] = 0.0
        for n, v in const.items():
            if n in SLOT_INDEX and isinstance(v, (int, float, np.integer, np.floating)):
                self.c.column[SLOT_INDEX[n]] = BJ_COL_CONST
                self.c.value[SLOT_INDEX[n]] = float(v)
        for col, n in enumerate(self.names):
            if n in SLOT_INDEX:
                self.c.column[SLOT_INDEX[n]] = col
        self.ignored = [n for n in self.names if n not in SLOT_INDEX and not n.startswith(self.EXTRA_PREFIXES)]

    @classmethod
    def from_dict(cls, params: Dict[str, float], ifos: Sequence[str] = (), nspcal: int = 0, nweights: int = 0):
        """Map for a single parameter dict (GWLikelihood.log_like(params))."""
        names = [n for n in PARAM_SLOTS if n in params]
        if nspcal or nweights:
            names += [n for n in params if n.startswith(cls.EXTRA_PREFIXES) and n != "weights"]
        m = cls(names, {}, ifos=ifos, nspcal=nspcal, nweights=nweights)
        x = np.array([float(params[n]) for n in names], dtype=np.float64)
        return m, x

    def present(self, name: str) -> bool:
        return self.c.column[SLOT_INDEX[name]] != BJ_COL_ABSENT


def make_config(device, approx, n_ifo, seglen, dets, marg_phi_ref=False, sincos="mufu",
                dd_sum=0.0, logZ_noise=0.0) -> bj_config:
    if approx not in APPROX_IDS:
        raise NotImplementedError(
            "approximant '%s' is not in the fused GPU path (TaylorF2 family only: %s)" % (approx, list(APPROX_IDS)))
    cfg = bj_config()
    cfg.abi_version = BJ_ABI_VERSION
    cfg.device = int(device)
    cfg.approx = APPROX_IDS[approx]
    cfg.n_ifo = int(n_ifo)
    cfg.marg_phi_ref = 1 if marg_phi_ref else 0
    cfg.sincos = SINCOS_MODES[sincos]
    cfg.seglen = float(seglen)
    cfg.dd_sum = float(dd_sum)
    cfg.logZ_noise = float(logZ_noise)
    for i, d in enumerate(dets):
        resp = np.asarray(d["response"], dtype=np.float64).reshape(9)
        loc = np.asarray(d["location"], dtype=np.float64).reshape(3)
        for j in range(9):
            cfg.det[i].response[j] = resp[j]
        for j in range(3):
            cfg.det[i].location[j] = loc[j]
        cfg.det[i].gmst_reference = float(d["gmst_reference"])
        cfg.det[i].reference_time = float(d["reference_time"])
        cfg.det[i].sday = float(d["sday"])
    return cfg


class Engine:
    """Owns one ``bj_ctx``.  ``kind`` in {'fullgrid', 'roq', 'relbin'}."""

    def __init__(self, handle, n_ifo: int, kind: str, n_bins: int = 0):
        self._lib = load_library()
        self._h = handle
        self.n_ifo = n_ifo
        self.kind = kind
        self.n_bins = n_bins
        self._focused = False

    # -- constructors --------------------------------------------------------------------------
    @classmethod
    def fullgrid(cls, cfg: bj_config, freqs, datas, psds, spcal_freqs=None, len_weights=None,
                 marg_time_shift=False, n_full=0, band_offset=0, compress=None) -> "Engine":
        """``compress``: None = library default (compressed frequency axis on), False = full grid only, True = on, or
        a dict ``{"mchirp_min": solar masses, "tau_max": seconds}`` giving the design bounds (include/bajes_b200.h:
        bj_extras.compress*).  Walkers outside the design are evaluated on the full grid in the same call."""
        lib = load_library()
        freqs = _as_f64(freqs)
        d_ri = [np.ascontiguousarray(np.asarray(d, dtype=np.complex128)).view(np.float64) for d in datas]
        p = [_as_f64(x) for x in psds]
        for a in d_ri:
            assert a.size == 2 * freqs.size
        for a in p:
            assert a.size == freqs.size
        h = C.c_void_p()
        ex = bj_extras()
        sf = lw = None
        if spcal_freqs is not None and len(spcal_freqs) > 0:
            sf = _as_f64(spcal_freqs)
            ex.nspcal, ex.spcal_freqs = sf.size, _ptr(sf)
        if len_weights is not None and len(len_weights) > 0:
            lw = np.ascontiguousarray(len_weights, dtype=np.int64)
            ex.nweights, ex.len_weights = lw.size, lw.ctypes.data_as(C.POINTER(C.c_int64))
        if marg_time_shift:
            ex.marg_time_shift, ex.n_full, ex.band_offset = 1, int(n_full), int(band_offset)
        if compress is not None:
            if isinstance(compress, dict):
                ex.compress = 1
                ex.compress_mchirp_min = float(compress.get("mchirp_min", 0.0))
                ex.compress_tau_max = float(compress.get("tau_max", 0.0))
            else:
                ex.compress = 1 if compress else -1
        _check(lib, lib.bj_create_ex(C.byref(h), C.byref(cfg), freqs.size, _ptr(freqs), _ptr_array(d_ri), _ptr_array(p),
                                     C.byref(ex)), "bj_create_ex")
        return cls(h, cfg.n_ifo, "fullgrid", freqs.size)

    @classmethod
    def roq(cls, cfg: bj_config, freqs_join, mask_omega, mask_psi, knots, coefs, psi_weights, tc_min, tc_max):
        lib = load_library()
        fj = _as_f64(freqs_join)
        mo = np.ascontiguousarray(mask_omega, dtype=np.int32)
        mp = np.ascontiguousarray(mask_psi, dtype=np.int32)
        kn = [_as_f64(k) for k in knots]
        cf = [np.ascontiguousarray(np.asarray(c, dtype=np.complex128)).view(np.float64) for c in coefs]
        ps = [_as_f64(p) for p in psi_weights]
        n_coef = np.asarray(coefs[0]).shape[0]
        assert np.asarray(coefs[0]).shape[1] == mo.size
        t = bj_roq_tables()
        t.n_join, t.freqs_join = fj.size, _ptr(fj)
        t.n_lin, t.mask_omega = mo.size, mo.ctypes.data_as(C.POINTER(C.c_int32))
        t.n_quad, t.mask_psi = mp.size, mp.ctypes.data_as(C.POINTER(C.c_int32))
        t.n_knots, t.n_coef = kn[0].size, n_coef
        ka, ca, pa = _ptr_array(kn), _ptr_array(cf), _ptr_array(ps)
        t.knots, t.coef_ri, t.psi_weights = ka, ca, pa
        t.tc_min, t.tc_max = float(tc_min), float(tc_max)
        h = C.c_void_p()
        _check(lib, lib.bj_roq_create(C.byref(h), C.byref(cfg), C.byref(t)), "bj_roq_create")
        return cls(h, cfg.n_ifo, "roq")

    @classmethod
    def relbin(cls, cfg: bj_config, fbin, h0_edges, sdat):
        lib = load_library()
        fb = _as_f64(fbin)
        h0 = [np.ascontiguousarray(np.asarray(h, dtype=np.complex128)).view(np.float64) for h in h0_edges]
        sd = np.ascontiguousarray(np.asarray(sdat, dtype=np.complex128)).view(np.float64)
        assert sd.size == 4 * cfg.n_ifo * (fb.size - 1) * 2
        h = C.c_void_p()
        _check(lib, lib.bj_relbin_create(C.byref(h), C.byref(cfg), fb.size, _ptr(fb), _ptr_array(h0), _ptr(sd)),
               "bj_relbin_create")
        return cls(h, cfg.n_ifo, "relbin")

    # -- calls ----------------------------------------------------------------------------------
    def loglike(self, x: np.ndarray, pmap: ParamMap) -> np.ndarray:
        """Host arrays in / out (copies are inside the call)."""
        x = _as_f64(x)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        out = np.empty(x.shape[0], dtype=np.float64)
        if x.shape[0] == 0:
            return out
        ex = C.byref(pmap.extra) if pmap.extra is not None else None
        _check(self._lib, self._lib.bj_loglike_ex(self._h, _ptr(x), x.shape[0], x.shape[1], C.byref(pmap.c), ex,
                                                  _ptr(out)),
               "bj_loglike")
        return out

    def loglike_device(self, x_ptr: int, n_walkers: int, ndim: int, pmap: ParamMap, out_ptr: int,
                       parts_ptr: int = 0, stream: int = 0) -> None:
        """Device pointers (ints), asynchronous on ``stream`` (a cudaStream_t handle as int)."""
        ex = C.byref(pmap.extra) if pmap.extra is not None else None
        _check(self._lib,
               self._lib.bj_loglike_device_ex(self._h, C.c_void_p(x_ptr), n_walkers, ndim, C.byref(pmap.c), ex,
                                              C.c_void_p(out_ptr), C.c_void_p(parts_ptr or None),
                                              C.c_void_p(stream or None)),
               "bj_loglike_device")

    def num_chunks(self) -> int:
        n = C.c_int32()
        _check(self._lib, self._lib.bj_num_chunks(self._h, C.byref(n)), "bj_num_chunks")
        return n.value

    def partial_device(self, x_ptr: int, n_walkers: int, ndim: int, pmap: ParamMap, chunk_begin: int, chunk_end: int,
                       sums_ptr: int, stream: int = 0) -> None:
        """Frequency-split: local partial sums [2 n_ifo][W] of the chunk range (device pointers)."""
        ex = C.byref(pmap.extra) if pmap.extra is not None else None
        _check(self._lib,
               self._lib.bj_partial_device(self._h, C.c_void_p(x_ptr), n_walkers, ndim, C.byref(pmap.c), ex,
                                           chunk_begin, chunk_end, C.c_void_p(sums_ptr), C.c_void_p(stream or None)),
               "bj_partial_device")

    def finish_device(self, sums_ptr: int, n_walkers: int, out_ptr: int, parts_ptr: int = 0, stream: int = 0) -> None:
        _check(self._lib,
               self._lib.bj_finish_device(self._h, C.c_void_p(sums_ptr), n_walkers, C.c_void_p(out_ptr),
                                          C.c_void_p(parts_ptr or None), C.c_void_p(stream or None)),
               "bj_finish_device")

    def project_waveform(self, x: np.ndarray, pmap: ParamMap) -> np.ndarray:
        x = _as_f64(x).reshape(-1)
        out = np.empty((self.n_ifo, self.n_bins), dtype=np.complex128)
        _check(self._lib, self._lib.bj_project_waveform(self._h, _ptr(x), x.size, C.byref(pmap.c),
                                                        out.view(np.float64).ctypes.data_as(_DP)),
               "bj_project_waveform")
        return out

    def last_kernel_ms(self):
        ms = (C.c_float * 3)()
        _check(self._lib, self._lib.bj_last_kernel_ms(self._h, ms), "bj_last_kernel_ms")
        return [ms[0], ms[1], ms[2]]

    def last_tile_ms(self) -> float:
        """Duration of the tile kernel of the last call alone (CUDA events around its launch).  Synchronises."""
        ms = C.c_float()
        _check(self._lib, self._lib.bj_last_tile_ms(self._h, C.byref(ms)), "bj_last_tile_ms")
        return float(ms.value)

    def table_info(self, table: int) -> Dict[str, int]:
        """table 0, 1 = compressed levels, 2 = full grid, 3 = focus: {"slots", "window_slots", "chunks", "window_chunks"}"""
        out = (C.c_int64 * 4)()
        _check(self._lib, self._lib.bj_table_info(self._h, int(table), out), "bj_table_info")
        return {"slots": int(out[0]), "window_slots": int(out[1]), "chunks": int(out[2]), "window_chunks": int(out[3])}

    def last_launch_info(self):
        info = (C.c_int32 * 8)()
        _check(self._lib, self._lib.bj_last_launch_info(self._h, info), "bj_last_launch_info")
        return list(info)

    def compress_info(self, level: int = 0) -> Dict[str, object]:
        """Compressed frequency axis `level` of this context: {"nodes", "kept_bins", "blocks", "nodes_per_block",
        "mchirp_min", "tau_max", "reach_rad", "margin", "n_bins"}; nodes == 0 when the level does not exist (the context
        then evaluates the full grid only if level == 0)."""
        info = (C.c_int32 * 4)()
        spec = (C.c_double * 4)()
        _check(self._lib, self._lib.bj_compress_info(self._h, int(level), info, spec), "bj_compress_info")
        return {"nodes": int(info[0]), "kept_bins": int(info[1]), "blocks": int(info[2]), "nodes_per_block": int(info[3]),
                "mchirp_min": float(spec[0]), "tau_max": float(spec[1]), "reach_rad": float(spec[2]),
                "margin": float(spec[3]), "n_bins": int(self.n_bins)}

    def last_levels(self, n_walkers: int) -> np.ndarray:
        """Table index of every walker of the last call (0, 1 = compressed levels, 2 = full grid).  Synchronises."""
        out = np.empty(int(n_walkers), dtype=np.int32)
        _check(self._lib, self._lib.bj_last_levels(self._h, int(n_walkers), out.ctypes.data_as(C.POINTER(C.c_int32))),
               "bj_last_levels")
        return out

    def last_level_counts(self):
        """Walkers of the last call per table: (level 0, level 1, full grid); with a focus table a fourth entry counts the
        walkers it evaluated.  Synchronises."""
        n = (C.c_int64 * 4)()
        _check(self._lib, self._lib.bj_last_level_counts(self._h, n), "bj_last_level_counts")
        out = tuple(int(v) for v in n)
        return out if self._focused else out[:3]

    def focus(self, x: np.ndarray, pmap: "ParamMap", slope_a: float = 0.0, slope_b: float = 0.0,
              slope_c: float = -1.0) -> Dict[str, int]:
        """Build the focus table around the point ``x`` (include/bajes_b200.h: bj_focus).  Returns its node counts."""
        x = _as_f64(x).reshape(-1)
        info = (C.c_int32 * 4)()
        _check(self._lib, self._lib.bj_focus(self._h, _ptr(x), x.size, C.byref(pmap.c), float(slope_a), float(slope_b),
                                             float(slope_c), info), "bj_focus")
        self._focused = True
        return {"nodes": int(info[0]), "kept_bins": int(info[1]), "blocks": int(info[2]), "nodes_per_block": int(info[3])}

    def unfocus(self) -> None:
        _check(self._lib, self._lib.bj_unfocus(self._h), "bj_unfocus")
        self._focused = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bj_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compress_axis(freqs, weights, nodes_per_run=0, reach_rad=0.0, mchirp_min=0.0, tau_max=0.0, margin=0.0):
    """Host-only mirror of the compressed frequency axis the full-grid context builds (capi.cu: build_compressed_axis):
    ``weights`` = complex [n_ifo, n_bins] -> (node_freqs [n], node_weights complex [n_ifo, n], info dict).  No GPU needed."""
    lib = load_library()
    f = _as_f64(freqs)
    w = [np.ascontiguousarray(np.asarray(x, dtype=np.complex128)).view(np.float64) for x in np.atleast_2d(weights)]
    out_f = np.empty(f.size, dtype=np.float64)
    out_w = [np.empty(2 * f.size, dtype=np.float64) for _ in w]
    n = C.c_int64()
    info = (C.c_int32 * 4)()
    _check(lib, lib.bj_compress_axis_host(f.size, _ptr(f), len(w), _ptr_array(w), int(nodes_per_run), float(reach_rad),
                                          float(mchirp_min), float(tau_max), float(margin), C.byref(n), _ptr(out_f),
                                          _ptr_array(out_w), info), "bj_compress_axis_host")
    m = int(n.value)
    return (out_f[:m].copy(), np.stack([x[:2 * m].view(np.complex128) for x in out_w]),
            {"nodes": int(info[0]), "kept_bins": int(info[1]), "blocks": int(info[2]), "nodes_per_block": int(info[3])})


def polarizations(approx: str, seglen: float, freqs, x: np.ndarray, pmap: ParamMap, centre_shift=True, device=0):
    lib = load_library()
    if approx not in APPROX_IDS:
        raise NotImplementedError("approximant '%s' is not in the fused GPU path" % approx)
    f = _as_f64(freqs)
    x = _as_f64(x).reshape(-1)
    hp = np.empty(f.size, dtype=np.complex128)
    hc = np.empty(f.size, dtype=np.complex128)
    _check(lib, lib.bj_polarizations(device, APPROX_IDS[approx], float(seglen), 1 if centre_shift else 0, f.size,
                                     _ptr(f), _ptr(x), x.size, C.byref(pmap.c),
                                     hp.view(np.float64).ctypes.data_as(_DP), hc.view(np.float64).ctypes.data_as(_DP)),
           "bj_polarizations")
    return hp, hc
