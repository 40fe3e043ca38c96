"""
TEST INFRASTRUCTURE.  Generates the committed golden vectors under ``tests/golden/`` (and the design-ASD
fixture under ``bajes_b200/data/``) by running the UNMODIFIED reference (``/root/reference``, imported under
``oracle/shims.py``) on the seeded synthetic injections of ``bajes_b200/synthetic.py``.

Run in the build container only (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py            # everything (a few minutes; the 128 s config dominates)
    python oracle/make_golden.py c1 roq     # a subset

The reference publishes no test vectors of its own: these files ARE the pin.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from bajes_b200 import synthetic as syn      # noqa: E402
from oracle import harness, shims            # noqa: E402

GOLDEN = os.path.join(_ROOT, "tests", "golden")
DATA = os.path.join(_ROOT, "bajes_b200", "data")


def write_design_asd():
    """The public LIGO-P1200087-v18 aLIGO / AdV design curves bajes bundles
    (bajes/pipe/data/gw/asd/design/*.txt, read by noise.py:46-75)."""
    shims.import_reference()
    from bajes.obs.gw.utils import read_asd
    fa, aa = read_asd("design", "H1")
    fv, av = read_asd("design", "V1")
    os.makedirs(DATA, exist_ok=True)
    np.savez_compressed(os.path.join(DATA, "design_asd.npz"), aLIGO_f=fa, aLIGO_asd=aa, AdV_f=fv, AdV_asd=av)
    print("design_asd.npz:", len(fa), len(fv))


def _walkers(cfg, n_narrow, n_wide, seed):
    xa, names = syn.draw_walkers(cfg, n_narrow, seed, "narrow")
    xb, _ = syn.draw_walkers(cfg, n_wide, seed + 1, "wide")
    X = np.concatenate([xa, xb])
    # a few deliberately unphysical rows (log_like.py:121-129 -> -inf)
    X[-1, names.index("distance")] = np.inf        # hp identically zero
    X[-2, names.index("mchirp")] = np.nan          # NaN waveform
    X[-3, names.index("cos_iota")] = 1.5           # arccos -> NaN
    return X, names


def _det_constants(dets, ifos):
    return dict(gmst_reference=np.array([dets[i].gmst_reference for i in ifos]),
                sday=np.array([dets[i].sday for i in ifos]),
                dd=np.array([np.real(dets[i]._dd) for i in ifos]))


def _parts(like, X, names, const, n):
    """per-detector (Re dh, Im dh, hh) of the first n rows, straight from Detector.compute_inner_products."""
    out = np.zeros((n, len(like.ifos), 3))
    for r in range(n):
        p = {k: float(v) for k, v in zip(names, X[r])}
        p.update(const)
        wave = like.wave.compute_hphc(p, roq=like.roq)
        for i, ifo in enumerate(like.ifos):
            dh, hh, dd = like.dets[ifo].compute_inner_products(wave, p, "freq", roq=like.roq)
            dh = dh if like.roq is not None else dh.sum()
            out[r, i] = (np.real(dh), np.imag(dh), np.real(hh))
    return out


def gen_c1():
    cfg = syn.CONFIGS["C1"]
    cl = harness.reference_classes()
    net = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = net
    const = syn.constants(cfg)
    X, names = _walkers(cfg, 64, 64, 11)
    out = {}
    for marg in (False, True):
        like = harness.reference_likelihood(cfg, marg_phi_ref=marg, net=net)
        out["logl_marg%d" % int(marg)] = harness.eval_batch(like, X, names, const)
        if not marg:
            out["parts"] = _parts(like, X, names, const, 16)
            out["logl_inj"] = like.log_like(syn.injection_params(cfg))
            out["logZ_noise"] = like.logZ_noise
    band = (f >= cfg["f_min"]) & (f <= cfg["f_max"])
    np.savez_compressed(os.path.join(GOLDEN, "c1_bbh.npz"), X=X, names=np.array(names),
                        data=np.stack([datas[i].freq_series for i in ifos]),
                        psd_band=np.stack([dets[i].psd for i in ifos]), freqs=f, band=band,
                        **_det_constants(dets, ifos), **out)
    print("c1: logL(inj) = %.9f, finite %d/%d" % (out["logl_inj"], np.isfinite(out["logl_marg0"]).sum(), len(X)))


def gen_c2(name, n_narrow, n_wide, approxes):
    cfg = syn.CONFIGS[name]
    cl = harness.reference_classes()
    t0 = time.time()
    net = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = net
    const = syn.constants(cfg)
    X, names = _walkers(cfg, n_narrow, n_wide, 21)
    out = {}
    for approx, nmax, marg in approxes:
        like = harness.reference_likelihood(cfg, approx=approx, marg_phi_ref=marg, net=net)
        key = "logl_%s_marg%d" % (approx, int(marg))
        out[key] = harness.eval_batch(like, X[:nmax], names, const)
        out["inj_" + key] = like.log_like(syn.injection_params(cfg))
        if approx == cfg["approx"] and not marg:
            out["parts"] = _parts(like, X, names, const, 8)
        print("  %s %s: %d points, %.1f s" % (name, key, nmax, time.time() - t0))
    # checks that let a test verify it re-synthesised the SAME data without storing 10^6 samples
    chk = np.array([[np.sum(np.abs(dets[i].data) ** 2), np.real(np.sum(dets[i].data * np.conj(dets[i].data[::-1])))]
                    for i in ifos])
    np.savez_compressed(os.path.join(GOLDEN, name.lower() + ".npz"), X=X, names=np.array(names),
                        data_check=chk, n_band=len(dets[ifos[0]].freqs), **_det_constants(dets, ifos), **out)


def gen_roq():
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=8.0, seed=3003)
    cl = harness.reference_classes()
    net = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = net
    const = syn.constants(cfg)
    basis = harness.make_roq_basis(cfg, n_lin=48, n_quad=24, seed=7)
    X, names = _walkers(cfg, 96, 96, 31)
    prior = harness.FakePrior(names, [[np.nanmin(x[np.isfinite(x)]), np.nanmax(x[np.isfinite(x)])] for x in X.T], {})
    prior.bounds[names.index("time_shift")] = [-0.05, 0.05]
    time_points = 300
    from bajes.pipe import initialize_roq
    with tempfile.TemporaryDirectory() as tmp:
        harness.write_jenpyroq(tmp, basis, cfg)
        roq = {ifo: {} for ifo in ifos}
        for ifo in ifos:
            mask = datas[ifo].mask
            ff = f[mask]
            psd = noises[ifo].interp_psd_pad(ff)
            fj, m_psi, m_om, psi, omega = initialize_roq(tmp, time_points, ff, datas[ifo].freq_series[mask], psd,
                                                         cfg["f_min"], cfg["f_max"], cfg["seglen"], prior,
                                                         cfg["approx"])
            roq[ifo]["omega_weights_interp"] = omega
            roq[ifo]["psi_weights"] = psi
        roq["freqs_join"], roq["mask_psi"], roq["mask_omega"] = fj, m_psi, m_om
    out = {}
    for marg in (False, True):
        like = harness.reference_likelihood(cfg, marg_phi_ref=marg, roq=roq, net=net)
        with np.errstate(all="ignore"):
            out["logl_marg%d" % int(marg)] = harness.eval_batch(like, X, names, const)
        if not marg:
            out["parts"] = _parts(like, X, names, const, 8)
    np.savez_compressed(os.path.join(GOLDEN, "roq_small.npz"), X=X, names=np.array(names), seglen=cfg["seglen"],
                        seed=cfg["seed"], time_points=time_points, t_bounds=np.array([-0.05, 0.05]),
                        psi=np.stack([roq[i]["psi_weights"] for i in ifos]),
                        omega_mid=np.stack([roq[i]["omega_weights_interp"](cfg["seglen"] / 2 + 1e-3) for i in ifos]),
                        **_det_constants(dets, ifos), **out)
    print("roq: finite %d/%d" % (np.isfinite(out["logl_marg0"]).sum(), len(X)))


def gen_extras():
    """C2_small with calibration envelopes and/or PSD weights (detector.py:517-531)."""
    cfg = syn.CONFIGS["C2_small"]
    cl = harness.reference_classes()
    const = syn.constants(cfg)
    out = {}
    Xb, names_b = _walkers(cfg, 48, 48, 51)
    for tag, nsp, nw, marg in (("spcal", 6, 0, False), ("weights", 0, 4, False), ("both", 5, 3, False), ("both", 5, 3, True)):
        ex = syn.extras_setup(cfg, nsp, nw)
        Xe, names_e = syn.extras_columns(cfg, len(Xb), 52, nsp, nw)
        X, names = np.concatenate([Xb, Xe], axis=1), names_b + names_e
        net = syn.build_network(cfg, cl)
        like = harness.reference_likelihood(cfg, marg_phi_ref=marg, net=net, extras=ex)
        key = "%s_marg%d" % (tag, int(marg))
        with np.errstate(all="ignore"):
            out["logl_" + key] = harness.eval_batch(like, X, names, const)
        out["X_" + key] = X
        out["names_" + key] = np.array(names)
        if ex["spcal_freqs"] is not None:
            out["spcal_freqs_" + key] = ex["spcal_freqs"]
        if ex["len_weights"] is not None:
            out["len_weights_" + key] = ex["len_weights"]
        print("  extras %s: finite %d/%d" % (key, np.isfinite(out["logl_" + key]).sum(), len(X)))
    dets = net[2]
    np.savez_compressed(os.path.join(GOLDEN, "extras_small.npz"), **_det_constants(dets, cfg["ifos"]), **out)


def gen_tmarg():
    """Time-shift marginalised likelihood (log_like.py:135-156) on C1 and the reduced C2."""
    shims.import_reference()
    from bajes.pipe.log_like import GWLikelihood
    out = {}
    for name, n_narrow, n_wide in (("C1", 48, 16), ("C2_small", 24, 8)):
        cfg = syn.CONFIGS[name]
        cl = harness.reference_classes()
        net = syn.build_network(cfg, cl)
        const = syn.constants(cfg)
        X, names = _walkers(cfg, n_narrow, n_wide, 61)
        for mphi in (False, True):
            like = GWLikelihood(net[0], net[1], net[2], net[3], net[4], cfg["srate"], cfg["seglen"], cfg["approx"],
                                marg_time_shift=True, marg_phi_ref=mphi)
            with np.errstate(all="ignore"):
                out["logl_%s_mphi%d" % (name, int(mphi))] = harness.eval_batch(like, X, names, const)
        out["X_" + name] = X
        out["names"] = np.array(names)
        out["gmst_" + name] = np.array([net[2][i].gmst_reference for i in cfg["ifos"]])
        out["sday_" + name] = np.array([net[2][i].sday for i in cfg["ifos"]])
        print("  tmarg %s: finite %d/%d" % (name, np.isfinite(out["logl_%s_mphi0" % name]).sum(), len(X)))
    np.savez_compressed(os.path.join(GOLDEN, "tmarg.npz"), **out)


def gen_rwalk():
    """dynesty random-walk proposal (dynesty.py:503-614): the unmodified ``sample_rwalk`` on seeded toy chains.
    dynesty itself is not installable offline; the proposal only reads ``dynesty.__version__`` (:511), so a stub
    module carrying a version string is registered before the import."""
    import types
    shims.import_reference()
    if "dynesty" not in sys.modules:
        stub = types.ModuleType("dynesty")
        stub.__version__ = "1.2.3"
        sys.modules["dynesty"] = stub
    import logging
    logging.disable(logging.WARNING)
    from bajes.inf.sampler.dynesty import BajesDynestyProposal
    from oracle import rwalk_cases as rc
    out = {}
    for group, g in rc.GROUPS.items():
        args, seeds = rc.make_args(group)
        prop = BajesDynestyProposal(g["maxmcmc"], walks=g["walks"], nact=g["nact"])
        rows = []
        for a, s in zip(args, seeds):
            np.random.seed(s)
            u, v, logl, ncall, blob = prop.propose(a)
            rows.append(np.concatenate([u, v, [logl, ncall, blob["accept"], blob["reject"], blob["fail"], blob["scale"]]]))
        out[group] = np.array(rows)
        r = out[group]
        print("  rwalk %-10s chains %2d  accept %s  fail %s" % (group, len(rows), r[:, -4].astype(int).tolist()[:8],
                                                                 r[:, -2].astype(int).tolist()[:8]))
    logging.disable(logging.NOTSET)
    np.savez_compressed(os.path.join(GOLDEN, "rwalk.npz"), **out)


RELBIN_CFG = dict(ifos=["H1", "L1", "V1"], seglen=8.0, srate=4096.0, f_min=20.0, f_max=2048.0,
                  approx="TaylorF2_3.5PN", source="bns", n_walkers=64, seed=4204)


def gen_relbin():
    """Relative-binning SET-UP pinned to reference code.  ``bajes/pipe/utils/binning.py`` cannot be imported as shipped
    (it imports ``bajes.pipe.likelihood`` and ``bajes.utils``, which do not exist: SURVEY.md F5); with two empty stub
    parent modules registered its two pure set-up functions ``setup_bins`` (:22-49) and ``compute_sdat`` (:52-101) run
    UNMODIFIED.  Inputs: the seeded synthetic network built with the reference classes, the fiducial waveform from the
    reference's own Waveform + Detector.project_fdwave on the masked band (the documented repair: h0 lives on the same
    axis as the detector data).  The per-point ``log_like`` of that module stays unrunnable, so logL parity for this
    likelihood remains unpinned; bins and summary data are not."""
    import types
    shims.import_reference()
    for name, attrs in (("bajes.pipe.likelihood", {"Likelihood": type("Likelihood", (), {})}),
                        ("bajes.utils", {"erase_init_wrapper": lambda f: f, "list_2_dict": lambda *a, **k: {}})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    import importlib
    binning = importlib.import_module("bajes.pipe.utils.binning")
    cfg = RELBIN_CFG
    cl = harness.reference_classes()
    ifos, datas, dets, noises, f = syn.build_network(cfg, cl)
    fid = syn.injection_params(cfg)
    mask = datas[ifos[0]].mask
    fband = np.asarray(f)[mask]
    wave = cl.Waveform(fband, cfg["srate"], cfg["seglen"], cfg["approx"])
    hphc = wave.compute_hphc(dict(fid))
    psds, sfts, h0s = [], [], []
    for ifo in ifos:
        dets[ifo].store_measurement(datas[ifo], noises[ifo])
        psds.append(np.asarray(dets[ifo].psd))
        sfts.append(np.asarray(dets[ifo].data))
        h0s.append(np.asarray(dets[ifo].project_fdwave(hphc, dict(fid), "freq")))
    out = {}
    for tag, chi, eps in (("default", 1.0, 0.5), ("fine", 1.0, 0.1)):
        nbin, fbin, fbin_ind = binning.setup_bins(fband, cfg["f_min"], cfg["f_max"], chi=chi, eps=eps)
        sdat = binning.compute_sdat(fband, fbin, fbin_ind, len(ifos), psds, sfts, h0s)
        out.update({"nbin_" + tag: nbin, "fbin_" + tag: fbin, "fbin_ind_" + tag: fbin_ind, "sdat_" + tag: sdat,
                    "chi_eps_" + tag: np.array([chi, eps])})
        print("  relbin %-8s chi=%g eps=%g  Nbin=%d  |A0| max %.3g" % (tag, chi, eps, nbin, np.abs(sdat[0]).max()))
    np.savez_compressed(os.path.join(GOLDEN, "relbin_setup.npz"), h0=np.array(h0s), **out, **_det_constants(dets, ifos))


def main(argv):
    what = set(argv) or {"asd", "c1", "c2_small", "c2", "roq", "extras", "tmarg", "rwalk", "relbin"}
    os.makedirs(GOLDEN, exist_ok=True)
    if "asd" in what:
        write_design_asd()
    if "c1" in what:
        gen_c1()
    if "c2_small" in what:
        gen_c2("C2_small", 128, 128,
               [("TaylorF2_3.5PN", 256, False), ("TaylorF2_3.5PN", 64, True), ("TaylorF2_5.5PN", 64, False),
                ("TaylorF2_5.5PN_7.5PNTides", 64, False), ("TaylorF2_5.5PN_3.5PNQM_7.5PNTides", 64, False),
                ("TaylorF2_5.5PN_7.5PNTides2020", 64, False)])
    if "c2" in what:
        gen_c2("C2", 128, 128, [("TaylorF2_3.5PN", 256, False), ("TaylorF2_5.5PN_7.5PNTides", 32, False)])
    if "roq" in what:
        gen_roq()
    if "extras" in what:
        gen_extras()
    if "tmarg" in what:
        gen_tmarg()
    if "rwalk" in what:
        gen_rwalk()
    if "relbin" in what:
        gen_relbin()


if __name__ == "__main__":
    main(sys.argv[1:])
