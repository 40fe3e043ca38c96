"""FP32-aware minimax polynomials for sin(pi z)/z and cos(pi z) on z in [-1/2, 1/2] (kernels.cuh: scaled_phasor).

What matters for the likelihood is the SMOOTH part of the phasor error (a function of the angle): consecutive
frequency bins have nearly equal phase, so smooth errors are correlated over hundreds of bins and do not average
down, whereas FP32 rounding noise does.  So the coefficients are chosen as float32 numbers such that the polynomial
they define -- evaluated EXACTLY -- is as close as possible to the true function (greedy "fpminimax": round the
leading coefficient, refit the rest, round the next, ...).  The leading sine coefficient is split hi + lo.
Prints coefficients, the smooth (exact-arithmetic) error, and the FP32-evaluation error statistics."""
import itertools
import numpy as np

M = 6001
T = np.cos(np.pi * (np.arange(M) + 0.5) / M)
X = 0.125 + 0.125 * T                     # z^2 in [0, 1/4]

def lawson(V, y, iters=60):
    w = np.ones(len(y))
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
        e = np.abs(V @ c - y)
        w = w * (e / e.mean() + 1e-3) ** 0.5
        w /= w.mean()
    return c

def fp_minimax(fun, n, split_lead=False):
    """coefficients c_0..c_{n-1} (in z^2), float32-representable, greedy rounding with refit."""
    y = fun(X)
    fixed = []
    for k in range(n):
        V = np.vander(X, n, increasing=True)
        yk = y - V[:, :k] @ np.array(fixed, dtype=np.float64) if k else y
        c = lawson(V[:, k:], yk)
        lead = c[0]
        if k == 0 and split_lead:
            hi = np.float32(lead)
            lo = np.float32(lead - np.float64(hi))
            fixed.append(np.float64(hi) + np.float64(lo))
            split = (hi, lo)
            continue
        cands = [np.float32(lead), np.nextafter(np.float32(lead), np.float32(np.inf)), np.nextafter(np.float32(lead), np.float32(-np.inf))]
        best = None
        for cand in cands:
            f2 = fixed + [np.float64(cand)]
            if k + 1 < n:
                yy = y - V[:, :k + 1] @ np.array(f2)
                cc = lawson(V[:, k + 1:], yy, iters=25)
                err = np.max(np.abs(V[:, k + 1:] @ cc - yy))
            else:
                err = np.max(np.abs(V[:, :k + 1] @ np.array(f2) - y))
            if best is None or err < best[0]:
                best = (err, cand)
        fixed.append(np.float64(best[1]))
    out = np.array(fixed)
    return (out, split) if split_lead else (out, None)

def f32fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

def report(ns, nc, split):
    cs, sp = fp_minimax(lambda x: np.where(x > 0, np.sin(np.pi * np.sqrt(x)) / np.sqrt(np.maximum(x, 1e-300)), np.pi), ns, split_lead=split)
    cc, _ = fp_minimax(lambda x: np.cos(np.pi * np.sqrt(x)), nc)
    z = np.linspace(-0.5, 0.5, 400001)
    z2 = z * z
    ps = np.polyval(cs[::-1], z2) * z
    pc = np.polyval(cc[::-1], z2)
    es, ec = ps - np.sin(np.pi * z), pc - np.cos(np.pi * z)
    amp = pc * np.cos(np.pi * z) + ps * np.sin(np.pi * z) - 1
    pha = ps * np.cos(np.pi * z) - pc * np.sin(np.pi * z)
    print("ns=%d nc=%d split=%s : SMOOTH err max sin %.2e cos %.2e | amp: mean %+.2e rms(zero-mean) %.2e | phase: mean %+.2e rms %.2e"
          % (ns, nc, split, np.abs(es).max(), np.abs(ec).max(), amp.mean(), (amp - amp.mean()).std(), pha.mean(), pha.std()))
    if sp:
        print("   sin lead hi %.9e lo %.9e" % (sp[0], sp[1]))
    print("   sin:", ", ".join("%.9ef" % v for v in cs.astype(np.float32)))
    print("   cos:", ", ".join("%.9ef" % v for v in cc.astype(np.float32)))
    return cs, cc, sp

for ns, nc, split in ((6, 6, False), (6, 6, True), (7, 7, False), (7, 7, True)):
    report(ns, nc, split)
