"""
ROQ basis construction on the GPU.  The reference ships only the run-time side of ROQ (loader
and weights, bajes/pipe/utils/roq.py:808-872); its own construction is commented-out code (roq.py:16-662).  This module
builds what that code was written to build -- linear and quadratic empirical interpolants in the JenpyROQ layout the loader
reads (roq.py:834-853) -- following it: training parameters = corners of the box + random draws, chirp mass drawn as
``min (max/min)^x`` (:35-37, :222-257); normalised ``h_plus`` of the approximant without the centre shift as training
functions (:298-323, :471-473); reduced-basis greedy under ``(a, b) = 4 df sum w conj(a) b``, ``w = 1/S`` (:537-612,
:654-662), in stages that keep only what the basis so far does not already represent (:325-362); DEIM nodes (:614-652).
Repair, stated: the commented ``inner_product`` keeps only the real part (:662); e and i e are then orthogonal, the greedy
admits both, and the complex DEIM sees linearly dependent columns (condition number 1e6 times worse on the test family,
tests/test_roq_basis.py) -- the Hermitian product of the cited algorithm (Antil et al., arXiv:1210.0577) is used.
Deviation, stated: the quadratic basis is trained on ``|h_plus|^2`` of the drawn parameters (what the ROQ rule needs,
detector.py:535) instead of products ``conj(e_i) e_j`` of linear training functions (:364-383).

The heavy steps run on the device through the C ABI (``bj_wavebank``, ``bj_rb_greedy``: csrc/roq_basis.cuh); torch holds the
device buffers.  DEIM works on an [n_freqs x m] matrix and stays on the host.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import engine as _engine


def draw_training_parameters(ranges, n, rng, names):
    """Rows of X [n_corners + n, len(names)]: the corners of the box of varied parameters first (roq.py:222-235), then n
    random points -- uniform, except ``mchirp = lo (hi/lo)^x`` (roq.py:35-37,237-251).  ``ranges[name]`` is a (lo, hi)
    pair or a constant."""
    varied = [k for k in names if np.ndim(ranges[k]) == 1]
    rows = []
    for corner in itertools.product(*[ranges[k] for k in varied]):
        p = {k: (corner[varied.index(k)] if k in varied else float(ranges[k])) for k in names}
        rows.append([p[k] for k in names])
    for _ in range(n):
        p = {}
        for k in names:
            if k not in varied:
                p[k] = float(ranges[k])
            else:
                lo, hi = ranges[k]
                x = rng.random()
                p[k] = lo * (hi / lo) ** x if k == "mchirp" else x * (hi - lo) + lo
        rows.append([p[k] for k in names])
    return np.ascontiguousarray(rows, dtype=np.float64)


def deim(V, quad_points):
    """DEIM nodes of the columns of V [M x m] (roq.py:614-652): (node frequencies, node indices)."""
    V = np.asarray(V)
    M, m = V.shape
    p = np.zeros(m, dtype=np.int64)
    U = np.zeros((M, m), dtype=V.dtype)
    U[:, 0] = V[:, 0]
    p[0] = int(np.argmax(np.abs(V[:, 0])))
    for i in range(1, m):
        c = np.linalg.solve(U[p[:i], :i], V[p[:i], i])
        r = V[:, i] - U[:, :i] @ c
        U[:, i] = r
        p[i] = int(np.argmax(np.abs(r)))
    return np.asarray(quad_points)[p], p


def _bank(device, approx, freqs, X, names, const, quadratic):
    import torch
    lib = _engine.load_library()
    pm = _engine.ParamMap(names, const)
    dev = torch.device("cuda", device)
    out = torch.empty((X.shape[0], len(freqs)), dtype=torch.complex128, device=dev)
    f = np.ascontiguousarray(freqs, dtype=np.float64)
    rc = lib.bj_wavebank(device, _engine.APPROX_IDS[approx], f.size, f.ctypes.data_as(C.POINTER(C.c_double)),
                         X.ctypes.data_as(C.POINTER(C.c_double)), X.shape[0], X.shape[1], C.byref(pm.c), int(quadratic),
                         C.c_void_p(out.data_ptr()))
    _engine._check(lib, rc, "bj_wavebank")
    return out


def _greedy(device, train, weights, df, tol, basis, n_prebuilt, max_basis):
    """train [n, M] complex128 (device, overwritten by residuals); basis [max_basis, M] complex128 (device)"""
    lib = _engine.load_library()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    picks = np.zeros(max_basis + 1, dtype=np.int32)
    sig = np.zeros(max_basis + 1, dtype=np.float64)
    n = C.c_int32()
    ms = C.c_float()
    rc = lib.bj_rb_greedy(device, train.shape[0], train.shape[1], C.c_void_p(train.data_ptr()),
                          w.ctypes.data_as(C.POINTER(C.c_double)), float(df), float(tol), int(n_prebuilt), int(max_basis),
                          C.c_void_p(basis.data_ptr()), picks.ctypes.data_as(C.POINTER(C.c_int32)),
                          sig.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n), C.byref(ms))
    _engine._check(lib, rc, "bj_rb_greedy")
    nb = int(n.value)
    return nb, picks[:nb - n_prebuilt], sig[:nb - n_prebuilt + 1], float(ms.value)


def reduced_basis(freqs, weights, df, approx, ranges, names, const, stages=(2048,), tol=1e-6, max_basis=512, quadratic=False,
                  seed=0, device=0):
    """Greedy reduced basis of the training family: returns (basis [m, M] complex128 (numpy), info).  Every stage draws a
    fresh training set, removes from it what the basis so far represents to ``tol`` (roq.py:259-296) and continues the
    greedy selection (:351-357)."""
    import torch
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", device)
    M = len(freqs)
    basis = torch.zeros((max_basis, M), dtype=torch.complex128, device=dev)
    w_dev = torch.as_tensor(np.asarray(weights, dtype=np.float64), device=dev)
    n_basis, info = 0, {"stages": [], "greedy_ms": 0.0, "params": []}
    for si, size in enumerate(stages):
        X = draw_training_parameters(ranges, int(size), rng, names) if si == 0 else \
            draw_training_parameters(ranges, int(size), rng, names)[2 ** sum(np.ndim(ranges[k]) == 1 for k in names):]
        bank = _bank(device, approx, freqs, X, names, const, quadratic)
        good = torch.isfinite(bank.real).all(dim=1)
        bank, X = bank[good].contiguous(), X[good.cpu().numpy()]
        nrm = torch.sqrt(4.0 * df * (w_dev * (bank.real ** 2 + bank.imag ** 2)).sum(dim=1))
        bank /= nrm[:, None]                                    # training functions are normalised (roq.py:315)
        n0 = n_basis
        n_basis, picks, sig, ms = _greedy(device, bank, weights, df, tol, basis, n0, max_basis)
        info["stages"].append({"training": int(bank.shape[0]), "basis_before": n0, "basis_after": n_basis,
                               "sigma_first": float(sig[0]) if len(sig) else None, "sigma_end": float(sig[-1]) if len(sig) else None,
                               "greedy_ms": ms})
        info["greedy_ms"] += ms
        info["params"] += [X[i] for i in picks]
        if n_basis >= max_basis:
            break
    info["sigmas"] = sig
    return basis[:n_basis].cpu().numpy(), info


def build_roq_basis(freqs, psd, seglen, approx, ranges, names, const=None, stages=(2048,), tol=1e-6, max_linear=512,
                    max_quadratic=256, seed=0, device=0):
    """Linear and quadratic empirical interpolants on the in-band axis ``freqs`` for a training box ``ranges``
    ({name: (lo, hi) | constant} over ``names``): dict(lin_f, quad_f, lin_b [M, n_lin] complex, quad_b [M, n_quad] real,
    info) in the layout of ``synthetic.write_jenpyroq`` / ``roq.load_jenpyroq`` (roq.py:834-853)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    weights = 1.0 / np.asarray(psd, dtype=np.float64)            # roq.py:31
    df = 1.0 / seglen
    const = dict(const or {})
    out = {}
    for tag, quad, mx in (("lin", False, max_linear), ("quad", True, max_quadratic)):
        basis, info = reduced_basis(freqs, weights, df, approx, ranges, names, const, stages=stages, tol=tol, max_basis=mx,
                                    quadratic=quad, seed=seed + (17 if quad else 0), device=device)
        V = basis.T                                             # [M, m]
        if quad:
            V = V.real.copy()                                   # |h|^2 is real: so is its basis
        nodes_f, p = deim(V, freqs)
        order = np.argsort(p)
        p, nodes_f = p[order], nodes_f[order]
        B = V @ np.linalg.inv(V[p, :])
        out[tag + "_f"] = nodes_f
        out[tag + "_b"] = np.ascontiguousarray(B if not quad else B.real)
        out[tag + "_info"] = info
    return out
