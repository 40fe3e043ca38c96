"""
Reduced-order-quadrature set-up mirroring bajes/pipe/utils/roq.py:808-872 and
bajes/pipe/__init__.py:661-716 (JenpyROQ bases, notation of arXiv:1604.08253).

Set-up only (once per detector): the linear weights omega_j(tc) are tabulated on a grid of
coalescence times and wrapped in a cubic spline exactly as the reference does with
``scipy.interpolate.interp1d(kind='cubic')``; the per-point evaluation (spline gather + two dots,
detector.py:534-537) runs in the ROQ CUDA kernel, which reads the FITTED B-spline (knots +
coefficients) rather than refitting.
"""
from __future__ import annotations

import os

import numpy as np


def compute_linear_weights(tc, df, interpolant, data, psd, freqs):
    """Eq. (9b) of arXiv:1604.08253 for one coalescence time (roq.py:808-815)."""
    vec = np.conj(data) / psd * np.exp(-2.0 * np.pi * 1j * tc * freqs)
    return 4.0 * df * (interpolant.T @ vec)


def linear_weights_matrix(tcs, df, interpolant, data, psd, freqs, block=256, device=None):
    """[N_L, N_tc] matrix of omega_j(tc) (roq.py:862-865), evaluated as blocked GEMMs.  With ``device`` (a CUDA
    device) the phase matrix and the complex128 GEMMs run there (set-up only; the same sums in another order:
    agreement with the host path ~1e-15 relative)."""
    base = np.conj(data) / psd
    if device is not None:
        import torch
        dev = torch.device(device)
        f_d = torch.as_tensor(np.asarray(freqs, dtype=np.float64), device=dev)
        base_d = torch.as_tensor(base, device=dev)
        bt_d = torch.as_tensor(np.ascontiguousarray(interpolant.T).astype(np.complex128), device=dev)
        t_d = torch.as_tensor(np.asarray(tcs, dtype=np.float64), device=dev)
        out_d = torch.empty((interpolant.shape[1], len(tcs)), dtype=torch.complex128, device=dev)
        blk = 1024
        for s in range(0, len(tcs), blk):
            ang = -2.0 * np.pi * torch.outer(f_d, t_d[s:s + blk])
            out_d[:, s:s + blk] = 4.0 * df * (bt_d @ (base_d[:, None] * torch.polar(torch.ones_like(ang), ang)))
        return out_d.cpu().numpy()
    out = np.empty((interpolant.shape[1], len(tcs)), dtype=np.complex128)
    bt = np.ascontiguousarray(interpolant.T)
    for s in range(0, len(tcs), block):
        t = np.asarray(tcs[s:s + block])
        ph = np.exp(-2.0 * np.pi * 1j * np.outer(freqs, t))          # [N, block]
        out[:, s:s + block] = 4.0 * df * (bt @ (base[:, None] * ph))
    return out


def load_jenpyroq(path):
    """JenpyROQ directory layout (roq.py:834-853)."""
    meta = None
    try:
        m = np.genfromtxt(os.path.join(path, "ROQ_data/ROQ_metadata.txt"), names=True)
        meta = {"f-min": float(m["fmin"]), "f-max": float(m["fmax"]), "seglen": float(m["seglen"])}
    except Exception:
        meta = None
    lin_f = np.load(os.path.join(path, "ROQ_data/linear/empirical_frequencies_linear.npy"))
    quad_f = np.load(os.path.join(path, "ROQ_data/quadratic/empirical_frequencies_quadratic.npy"))
    lin_b = np.load(os.path.join(path, "ROQ_data/linear/basis_interpolant_linear.npy"))
    quad_b = np.load(os.path.join(path, "ROQ_data/quadratic/basis_interpolant_quadratic.npy"))
    return meta, lin_f, quad_f, lin_b, quad_b


def initialise_roq_for_inference(path, approx, f_min, f_max, df, freqs_full, psd, data_f, tcs, device=None):
    """roq.py:817-872: returns linear / quadratic node frequencies, the cubic interpolant of the linear
    weights over ``tcs`` and the quadratic weights."""
    from scipy.interpolate import interp1d
    meta, lin_f, quad_f, lin_b, quad_b = load_jenpyroq(path)
    if meta is not None:
        if not (f_min >= meta["f-min"] and f_max <= meta["f-max"] and df >= 1.0 / meta["seglen"]):
            # the reference raises inside a bare try/except and then falls into the legacy branch
            meta = None
    if meta is None:
        data_f, psd, freqs_full = data_f[:-1], psd[:-1], freqs_full[:-1]   # legacy PyROQ layout (roq.py:844-849)
    matrix = linear_weights_matrix(tcs, df, lin_b, data_f, psd, freqs_full, device=device)
    omega = interp1d(tcs, matrix, kind="cubic", bounds_error=False, fill_value=-np.inf)
    psi = 4.0 * df * quad_b.T @ (1.0 / psd)
    return lin_f, quad_f, omega, psi


def initialize_roq(roq_path, time_points, freqs_full, data_f, psd, f_min, f_max, seglen, prior, approx, device=None):
    """bajes/pipe/__init__.py:661-716: tc grid from the time_shift prior (+-30 ms), weights, joined node
    axis and the index masks."""
    if "time_shift" not in prior.const:
        idx = list(prior.names).index("time_shift")
        t_low, t_high = prior.bounds[idx][0], prior.bounds[idx][1]
    else:
        t_low = t_high = prior.const["time_shift"]
    tcs = np.linspace(seglen / 2.0 + t_low - 0.030, seglen / 2.0 + t_high + 0.030, time_points)
    lin_f, quad_f, omega, psi = initialise_roq_for_inference(roq_path, approx, f_min, f_max, 1.0 / seglen,
                                                             freqs_full, psd, data_f, tcs, device=device)
    joined = np.sort(np.array(list(set(np.concatenate((quad_f, lin_f))))))
    mask_lin = [i for i in range(len(joined)) if joined[i] in lin_f]
    mask_qua = [i for i in range(len(joined)) if joined[i] in quad_f]
    return joined, mask_qua, mask_lin, psi, omega


def build_roq_dict(roq_path, time_points, ifos, datas, noises, freqs, f_min, f_max, seglen, prior, approx, device=None):
    """The ``roq`` argument of GWLikelihood as assembled in bajes/pipe/__init__.py:753-783
    (PSD WITHOUT the window factor, :763).  ``device``: evaluate the weight matrices on that CUDA device."""
    roq = {ifo: {} for ifo in ifos}
    joined = m_psi = m_om = None
    for ifo in ifos:
        mask = datas[ifo].mask
        f_full = np.asarray(freqs)[mask]
        d = datas[ifo].freq_series[mask]
        psd = noises[ifo].interp_psd_pad(f_full)
        joined, m_psi, m_om, psi, omega = initialize_roq(roq_path, time_points, f_full, d, psd, f_min, f_max,
                                                         seglen, prior, approx, device=device)
        roq[ifo]["omega_weights_interp"] = omega
        roq[ifo]["psi_weights"] = psi
    roq["freqs_join"] = joined
    roq["mask_psi"] = m_psi
    roq["mask_omega"] = m_om
    return roq


def spline_tables(interp):
    """(knots, coefficients [n_coef, N_L] complex, tc_min, tc_max) of the fitted cubic B-spline inside a
    ``scipy.interpolate.interp1d(kind='cubic')`` (or of a ``BSpline`` passed directly)."""
    spl = getattr(interp, "_spline", interp)
    if getattr(spl, "k", None) != 3:
        raise ValueError("ROQ linear weights must be a cubic spline (roq.py:866)")
    t = np.ascontiguousarray(spl.t, dtype=np.float64)
    c = np.asarray(spl.c, dtype=np.complex128)
    if c.ndim != 2:
        c = c.reshape(c.shape[0], -1)
    n = c.shape[0]
    return t, np.ascontiguousarray(c), float(t[3]), float(t[n])
