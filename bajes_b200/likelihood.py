"""
Drop-in ``GWLikelihood`` (bajes/pipe/log_like.py:21-183) backed by the fused sm_100a kernels, plus the
batched entry points the sampler adapters call, and the ``Likelihood`` / ``Posterior`` wrappers of
bajes/inf/likelihood.py with batched counterparts.

Per-point ``log_like(params)`` keeps the reference contract (dict in, float out, ``-inf`` for
unphysical points) and is evaluated on the GPU as a batch of one.  There is no CPU code path for
the likelihood: without the CUDA library / a B200 the constructor raises ``EngineUnavailable``.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Sequence

import numpy as np

from . import engine as _engine
from .waveform import Waveform

logger = logging.getLogger(__name__)


class Likelihood(object):
    """Generic likelihood wrapper (bajes/inf/likelihood.py:9-33)."""

    def __init__(self, func=None, args=[], kwargs={}):
        self.logZ_noise = 0.0
        self._func = func
        self._args = args
        self._kwargs = kwargs

    def log_like(self, x):
        if self._func is None:
            raise RuntimeError("Log-likelihood method has been called without specifying the likelihood form. "
                               "Please implement your Likelihood.")
        logl = self._func(x, *self._args, **self._kwargs)
        if np.isnan(logl):
            raise RuntimeError("Likelihood method returned NaN for the set of parameters: {}.".format(x))
        return logl


class GWLikelihood(Likelihood):
    """Frequency-domain Gaussian likelihood  Re(d|h) - (h|h)/2  over a detector network."""

    def __init__(self, ifos, datas, dets, noises, freqs, srate, seglen, approx,
                 nspcal=0, spcal_freqs=None, nweights=0, len_weights=None,
                 marg_phi_ref=False, marg_time_shift=False, roq=None,
                 device=0, sincos="mufu", compress=None, **kwargs):
        """``compress`` (no counterpart in the reference): None = library default (the compressed frequency axis is
        used where it pays), False = evaluate every bin, or {"mchirp_min": .., "tau_max": ..} = the design bounds of
        the compression (chirp mass in solar masses, |geocentric delay + time_shift| in s; defaults 0.8 / 0.11).  Points
        outside the design are evaluated on the full grid in the same call: the option changes speed, never results
        beyond the parity tolerance."""
        super(GWLikelihood, self).__init__()
        self.compress = compress
        self.ifos = list(ifos)
        self.dets = dets
        self.roq = roq
        self.nspcal = nspcal
        self.nweights = nweights
        self.marg_phi_ref = marg_phi_ref
        self.marg_time_shift = marg_time_shift
        self.device = device
        self.sincos = sincos
        self.approx = approx
        self.srate = srate
        self.seglen = seglen

        if self.roq is not None and self.marg_time_shift:
            raise AttributeError("Time-shift marginalization has not been implemented with the ROQ approximation.")
        if self.marg_time_shift and (nspcal or nweights):
            raise NotImplementedError("time-shift marginalisation together with calibration envelopes / PSD weights is "
                                      "not implemented on the GPU; there is no CPU fallback")
        if (nspcal or nweights) and roq is not None:
            raise AttributeError("The ROQ approximation has not yet been extended to incorporate calibration "
                                 "uncertainties.")          # bajes/pipe/__init__.py:737-739
        if nspcal > _engine.BJ_MAX_SPCAL or nweights > _engine.BJ_MAX_WEIGHTS:
            raise NotImplementedError("at most %d calibration nodes / %d PSD-weight bands"
                                      % (_engine.BJ_MAX_SPCAL, _engine.BJ_MAX_WEIGHTS))
        self.spcal_freqs = None if not nspcal else np.asarray(spcal_freqs, dtype=float)
        self.len_weights = None if not nweights else np.asarray(len_weights, dtype=np.int64)
        # networks of more than BJ_MAX_IFO detectors (the reference loops over any number, log_like.py:57-96: ET1-3 + CE
        # ...) are evaluated in passes of at most BJ_MAX_IFO detectors each; <d|h> and <h|h> are additive over detectors
        self._groups = [self.ifos[k:k + _engine.BJ_MAX_IFO] for k in range(0, len(self.ifos), _engine.BJ_MAX_IFO)]
        if len(self._groups) > 1 and self.marg_time_shift:
            raise NotImplementedError("time-shift marginalisation with more than %d detectors" % _engine.BJ_MAX_IFO)

        # consistency checks of log_like.py:57-96
        n_freqs = f_min_check = f_max_check = None
        mask = None
        for ifo in self.ifos:
            self.dets[ifo].store_measurement(datas[ifo], noises[ifo], nspcal, spcal_freqs, nweights, len_weights)
            self.dets[ifo]._owner = self
            if f_min_check is None:
                f_min_check = datas[ifo].f_min
            elif datas[ifo].f_min != f_min_check:
                raise ValueError("Input f_min of data and model do not match in detector {}.".format(ifo))
            if f_max_check is None:
                f_max_check = datas[ifo].f_max
            elif datas[ifo].f_max != f_max_check:
                raise ValueError("Input f_max of data and model do not match in detector {}.".format(ifo))
            if n_freqs is None:
                n_freqs = len(datas[ifo].freqs)
            elif len(datas[ifo].freqs) != n_freqs:
                raise ValueError("Number of data samples does not match in detector {}.".format(ifo))
            if datas[ifo].seglen != seglen:
                raise ValueError("Input seglen of data and model do not match in detector {}.".format(ifo))
            self.logZ_noise += -0.5 * self.dets[ifo]._dd
            self.Nfr = n_freqs
            mask = datas[ifo].mask
        idx = np.asarray(mask[0] if isinstance(mask, tuple) else mask)
        self._band_offset = int(idx[0]) if len(idx) else 0
        if self.marg_time_shift and len(idx) and not np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx))):
            raise NotImplementedError("time-shift marginalisation needs a contiguous in-band mask")

        freqs = np.asarray(freqs)
        if self.roq is not None:
            self.wave = Waveform(self.roq["freqs_join"], srate, seglen, approx, device=device)
        else:
            self.wave = Waveform(freqs[mask], srate, seglen, approx, device=device)
        self._engine = None
        self._build_engine()

    # ---- engine life cycle ------------------------------------------------------------------------
    def _config(self, ifos=None):
        ifos = self.ifos if ifos is None else ifos
        dd_sum = 0.0
        for ifo in ifos:
            dd_sum += np.real(self.dets[ifo]._dd)
        single = len(self._groups) == 1
        return _engine.make_config(self.device, self.approx, len(ifos), self.seglen,
                                   [self.dets[i].constants() for i in ifos],
                                   marg_phi_ref=self.marg_phi_ref and single, sincos=self.sincos,
                                   dd_sum=dd_sum, logZ_noise=self.logZ_noise if single else -0.5 * dd_sum)

    def _build_one(self, ifos):
        cfg = self._config(ifos)
        if self.roq is None:
            d0 = self.dets[ifos[0]]
            return _engine.Engine.fullgrid(cfg, d0.freqs,
                                           [self.dets[i].data for i in ifos],
                                           [self.dets[i].psd for i in ifos],
                                           spcal_freqs=self.spcal_freqs, len_weights=self.len_weights,
                                           marg_time_shift=self.marg_time_shift, n_full=self.Nfr,
                                           band_offset=self._band_offset, compress=getattr(self, "compress", None))
        from .roq import spline_tables
        knots, coefs, tc_min, tc_max = [], [], None, None
        for ifo in ifos:
            t, c, lo, hi = spline_tables(self.roq[ifo]["omega_weights_interp"])
            knots.append(t)
            coefs.append(c)
            tc_min = lo if tc_min is None else max(tc_min, lo)
            tc_max = hi if tc_max is None else min(tc_max, hi)
        return _engine.Engine.roq(cfg, self.roq["freqs_join"], self.roq["mask_omega"],
                                  self.roq["mask_psi"], knots, coefs,
                                  [self.roq[i]["psi_weights"] for i in ifos], tc_min, tc_max)

    def _build_engine(self):
        self._engines = [self._build_one(g) for g in self._groups]
        self._engine = self._engines[0]

    @property
    def engines(self):
        if self._engine is None:        # after unpickling: re-upload lazily
            self._build_engine()
        return self._engines

    @property
    def engine(self):
        """The CUDA context of a network of at most BJ_MAX_IFO detectors (larger networks own several: ``engines``)."""
        engines = self.engines
        if len(engines) > 1:
            raise NotImplementedError("this network is evaluated in %d passes: use log_like_batch / log_like / "
                                      "inner_products, or the per-pass contexts in .engines" % len(engines))
        return engines[0]

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None         # device handles never travel through pickle
        state["_engines"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        for ifo in self.ifos:
            self.dets[ifo]._owner = self

    # ---- evaluation -------------------------------------------------------------------------------
    def param_map(self, names: Sequence[str], const: Optional[Dict[str, float]] = None) -> _engine.ParamMap:
        return _engine.ParamMap(names, const, ifos=self.ifos, nspcal=self.nspcal, nweights=self.nweights)

    def _point_map(self, params):
        return _engine.ParamMap.from_dict(_physical_subset(params), ifos=self.ifos, nspcal=self.nspcal,
                                          nweights=self.nweights)

    def log_like_batch(self, X, names: Sequence[str], const: Optional[Dict[str, float]] = None) -> np.ndarray:
        """``[log_like({**dict(zip(names, x)), **const}) for x in X]`` in one GPU call.
        X: float64 [W, ndim] host array; returns float64 [W]."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != len(names):
            raise ValueError("X must be [n_points, %d]" % len(names))
        if len(self._groups) == 1:
            return self.engine.loglike(X, self.param_map(names, const))
        return self._multi_pass(X, [_engine.ParamMap(names, const, ifos=g, nspcal=self.nspcal, nweights=self.nweights)
                                    for g in self._groups])[0]

    def focus(self, params, names=None, const=None, slope_a=0.0, slope_b=0.0, slope_c=-1.0):
        """Tell the likelihood where the sampler is (no counterpart in the reference): builds the FOCUS table around one
        point -- ``params`` a dict, or a row of values with ``names`` / ``const`` as in ``log_like_batch`` -- so that points
        whose phase evolution stays within ``slope_a (f/f_min)^(-5/3) + slope_b + slope_c (f/f_max)^(2/3)`` seconds of the
        fiducial's (defaults 0.25 / 2e-3 / 0.03: the posterior of a long signal, tides free) are evaluated on a few hundred
        all-FP64 nodes.  Everything else is routed as
        before; results never depend on the focus beyond the parity tolerance.  Returns the node counts."""
        if len(self._groups) != 1:
            raise NotImplementedError("focus tables for networks of more than %d detectors" % _engine.BJ_MAX_IFO)
        if isinstance(params, dict):
            sub = _physical_subset(params)
            pmap, x = _engine.ParamMap.from_dict(sub)
        else:
            x = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
            pmap = self.param_map(list(names), const)
        return self.engine.focus(x, pmap, slope_a, slope_b, slope_c)

    def unfocus(self):
        self.engine.unfocus()

    def _multi_pass(self, X, pmaps):
        """Networks of more than BJ_MAX_IFO detectors: one pass per group of detectors on the same device-resident
        batch; per-pass logL (without phase marginalisation) and per-detector <d|h>, <h|h> are combined as
        log_like.py:158-181 does.  Returns (logL [W], parts [W, n_ifo, 3])."""
        import torch
        dev = torch.device("cuda", self.device)
        W = X.shape[0]
        parts_all = np.empty((W, len(self.ifos), 3))
        if W == 0:
            return np.empty(0), parts_all
        xd = torch.from_numpy(X).to(dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        outs, parts = [], []
        for g, eng, pm in zip(self._groups, self.engines, pmaps):
            o = torch.empty(W, dtype=torch.float64, device=dev)
            p = torch.empty((W, len(g), 3), dtype=torch.float64, device=dev)
            eng.loglike_device(xd.data_ptr(), W, X.shape[1], pm, o.data_ptr(), p.data_ptr(), st)
            outs.append(o)
            parts.append(p)
        logl = torch.stack(outs).sum(dim=0).cpu().numpy()
        k = 0
        for g, p in zip(self._groups, parts):
            parts_all[:, k:k + len(g)] = p.cpu().numpy()
            k += len(g)
        if self.marg_phi_ref:
            from scipy.special import i0e
            dh = parts_all[:, :, 0].sum(axis=1) + 1j * parts_all[:, :, 1].sum(axis=1)
            with np.errstate(invalid="ignore"):
                a = np.abs(dh)
                logl = logl - np.real(dh) + (np.log(i0e(a)) + a)            # log_like.py:176-177
            logl = np.where(np.isnan(logl), -np.inf, logl)
        return logl, parts_all

    def log_like(self, params):
        """Reference per-point contract (log_like.py:107-183): dict -> float."""
        if len(self._groups) > 1:
            x, pmaps = self._point_maps(params)
            return float(self._multi_pass(x.reshape(1, -1), pmaps)[0][0])
        pmap, x = self._point_map(params)
        return float(self.engine.loglike(x.reshape(1, -1), pmap)[0])

    def _point_maps(self, params):
        """one parameter row + a map per pass (the extras of a pass are looked up under its own detector names)"""
        sub = _physical_subset(params)
        names = [n for n in _engine.PARAM_SLOTS if n in sub]
        names += [n for n in sub if n.startswith(_engine.ParamMap.EXTRA_PREFIXES) and n != "weights"]
        x = np.array([float(sub[n]) for n in names], dtype=np.float64)
        return x, [_engine.ParamMap(names, {}, ifos=g, nspcal=self.nspcal, nweights=self.nweights) for g in self._groups]

    def _parts(self, params):
        """per-detector Re<d|h>, Im<d|h>, <h|h> of one point via the device entry."""
        import torch
        if len(self._groups) > 1:
            x, pmaps = self._point_maps(params)
            logl, parts = self._multi_pass(x.reshape(1, -1), pmaps)
            return parts[0], float(logl[0])
        pmap, x = self._point_map(params)
        dev = torch.device("cuda", self.device)
        xd = torch.from_numpy(x.reshape(1, -1)).to(dev)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        parts = torch.empty((1, len(self.ifos), 3), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self.engine.loglike_device(xd.data_ptr(), 1, x.size, pmap, out.data_ptr(), parts.data_ptr(), st)
        torch.cuda.synchronize(dev)
        return parts.cpu().numpy()[0], float(out.item())

    def inner_products(self, params):
        """{ifo: (<d|h> complex, <h|h> float)} for one point (Detector.compute_inner_products)."""
        parts, _ = self._parts(params)
        return {ifo: (complex(parts[i, 0], parts[i, 1]), float(parts[i, 2])) for i, ifo in enumerate(self.ifos)}

    def project_fdwave(self, params):
        """{ifo: projected strain on the in-band axis} for one point (Detector.project_fdwave)."""
        if self.roq is not None:
            raise NotImplementedError("projected strain on the full axis is not defined for an ROQ likelihood")
        sub = _physical_subset(params)
        pmap, x = _engine.ParamMap.from_dict(sub)
        out = {}
        for g, eng in zip(self._groups, self.engines):
            h = eng.project_waveform(x, pmap)
            out.update({ifo: h[i] for i, ifo in enumerate(g)})
        return out


def _physical_subset(params):
    """Keep the slots the device prologue understands, in the ORIGINAL parametrisation: when the caller's
    dict already carries the derived entries compute_hphc adds (mtot from mchirp, iota from cos_iota,
    cartesian from spherical spins) the primary ones win, as in waveform.py:334-377."""
    for k in params:
        if k.startswith(_engine.ParamMap.UNSUPPORTED_PREFIXES):
            raise NotImplementedError("parameter '%s' is outside the fused GPU likelihood; no CPU fallback" % k)
    sub = {k: v for k, v in params.items()
           if k in _engine.SLOT_INDEX or (k.startswith(_engine.ParamMap.EXTRA_PREFIXES) and k != "weights")}
    if "mchirp" in sub:
        sub.pop("mtot", None)
    if "cos_iota" in sub:
        sub.pop("iota", None)
    if "s1" in sub and "tilt1" in sub and "phi_1l" in sub:
        for k in ("s1x", "s1y", "s1z"):
            sub.pop(k, None)
    if "s2" in sub and "tilt2" in sub and "phi_2l" in sub:
        for k in ("s2x", "s2y", "s2z"):
            sub.pop(k, None)
    return sub


class Posterior(object):
    """bajes/inf/likelihood.py:35-80 plus batched counterparts with identical semantics row by row."""

    def __init__(self, like, prior):
        self.like = like
        self.prior = prior

    # -- reference per-point API --------------------------------------------------------------------
    def prior_transform(self, u):
        return self.prior.prior_transform(u)

    def prior_transform_batch(self, U):
        if hasattr(self.prior, "prior_transform_batch"):
            return np.asarray(self.prior.prior_transform_batch(U), dtype=np.float64)
        return np.stack([np.asarray(self.prior.prior_transform(u), dtype=np.float64) for u in U])

    def _fill(self, x):
        return self.prior.this_sample({k: v for k, v in zip(self.prior.names, x)})

    def log_like(self, x):
        return self.like.log_like(self._fill(x))

    def log_prior(self, x):
        return self.prior.log_prior(x) if self.prior.in_bounds(x) else -np.inf

    def log_post(self, x):
        if self.prior.in_bounds(x):
            return self.prior.log_prior(x) + self.like.log_like(self._fill(x))
        return -np.inf

    def log_likeprior(self, x):
        if self.prior.in_bounds(x):
            return self.like.log_like(self._fill(x)), self.prior.log_prior(x)
        return 0, -np.inf

    # -- batched API ------------------------------------------------------------------------------------
    def _batch_inputs(self, X):
        """Columns = sampled parameters (+ prior 'variables' evaluated on the host), constants separate."""
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        names = list(self.prior.names)
        v_names = list(getattr(self.prior, "v_names", []) or [])
        if v_names:
            extra = np.empty((X.shape[0], len(v_names)))
            for r in range(X.shape[0]):
                p = {k: v for k, v in zip(names, X[r])}
                for j, (f, kw) in enumerate(zip(self.prior.v_funcs, self.prior.v_kwargs)):
                    extra[r, j] = f(**p, **kw)
            X = np.concatenate([X, extra], axis=1)
            names = names + v_names
        const = {k: v for k, v in getattr(self.prior, "const", {}).items()}
        return X, names, const

    def log_like_batch(self, X):
        Xf, names, const = self._batch_inputs(X)
        return self.like.log_like_batch(Xf, names, const)

    def _prior_batch(self, X):
        """(in-bounds mask, log-prior of the in-bounds rows); uses the prior's vectorised methods when it has them
        (a bajes ``Prior`` is evaluated row by row, as bajes itself does)."""
        if hasattr(self.prior, "in_bounds_batch"):
            inb = np.asarray(self.prior.in_bounds_batch(X), dtype=bool)
        else:
            inb = np.array([bool(self.prior.in_bounds(x)) for x in X])
        if hasattr(self.prior, "log_prior_batch"):
            lnp = np.asarray(self.prior.log_prior_batch(X[inb]), dtype=float)
        else:
            lnp = np.array([self.prior.log_prior(x) for x in X[inb]], dtype=float)
        return inb, lnp

    def log_post_batch(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        inb, lnp = self._prior_batch(X)
        out = np.full(X.shape[0], -np.inf)
        if inb.any():
            out[inb] = lnp + self.log_like_batch(X[inb])
        return out

    def log_likeprior_batch(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        inb, lnp_in = self._prior_batch(X)
        lnl = np.zeros(X.shape[0])
        lnp = np.full(X.shape[0], -np.inf)
        if inb.any():
            lnp[inb] = lnp_in
            lnl[inb] = self.log_like_batch(X[inb])
        return lnl, lnp
