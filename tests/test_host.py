"""
CPU tests of the host side: set-up classes against the oracle port, parameter packing, the C-ABI
library (loads, exports every declared symbol, fails loudly without a GPU), pool adapter, pickling.
No likelihood value is computed on the CPU by the product -- there is no such code path.
"""
import os
import pickle
import re

import numpy as np
import pytest

from bajes_b200 import engine, synthetic as syn
from bajes_b200.detector import Detector
from bajes_b200.noise import Noise, design_asd
from bajes_b200.pool import BatchedPool
from bajes_b200.series import Series
from oracle import bajes_port as port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        engine.device_info(0)
        return False
    except engine.EngineUnavailable:
        return True


# ---- C ABI -----------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    header = open(os.path.join(ROOT, "include", "bajes_b200.h")).read()
    declared = set(re.findall(r"\b(bj_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bj_abi_version() == engine.BJ_ABI_VERSION


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(engine.bj_param_map) == engine.BJ_NUM_PARAMS * 4 + engine.BJ_NUM_PARAMS * 8
    assert C.sizeof(engine.bj_detector) == 15 * 8
    assert C.sizeof(engine.bj_config) == 6 * 4 + 3 * 8 + 3 * 15 * 8
    header = open(os.path.join(ROOT, "include", "bajes_b200.h")).read()
    slots = re.search(r"typedef enum \{\s*BJ_P_MCHIRP(.*?)BJ_NUM_PARAMS", header, re.S).group(0)
    names = [n.lower()[5:] for n in re.findall(r"BJ_P_[A-Z0-9_]+", slots)]
    assert names == list(engine.PARAM_SLOTS)


def test_no_silent_cpu_fallback():
    """Without a GPU the product must raise, never compute on the CPU."""
    if not _no_gpu():
        pytest.skip("a GPU is visible")
    cfg = syn.CONFIGS["C1"]
    cl = syn.product_classes()
    with pytest.raises(engine.EngineUnavailable):
        cl.Waveform(np.linspace(20, 100, 10), 4096., 4., cfg["approx"]).compute_hphc(syn.injection_params(cfg))
    with pytest.raises(engine.EngineUnavailable):
        engine.measure_peak("dfma")
    # the package never imports the oracle
    import bajes_b200
    pkg_dir = os.path.dirname(bajes_b200.__file__)
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_unsupported_features_raise():
    with pytest.raises(NotImplementedError):
        engine.ParamMap(["mchirp", "q", "spcal_amp0_H1"], {})          # likelihood built without calibration nodes
    with pytest.raises(NotImplementedError):
        engine.ParamMap(["mchirp", "q", "eos_logp1"], {})
    with pytest.raises(NotImplementedError):
        engine.make_config(0, "TEOBResumS", 2, 4.0, [])
    with pytest.raises(KeyError):
        engine.ParamMap(["mchirp", "q"], {}, ifos=["H1"], nspcal=2)     # calibration parameters missing


def test_param_map_extras_columns_and_constants():
    names = ["mchirp", "spcal_amp0_H1", "spcal_phi1_L1", "weight0_H1"]
    const = {"spcal_amp1_H1": 0.01, "spcal_phi0_H1": 0.0, "spcal_phi1_H1": 0.02, "spcal_amp0_L1": 0.0,
             "spcal_amp1_L1": 0.0, "spcal_phi0_L1": 0.0, "weight0_L1": 1.5, "weight1_H1": 1.0, "weight1_L1": 0.9}
    m = engine.ParamMap(names, const, ifos=["H1", "L1"], nspcal=2, nweights=2)
    e = m.extra
    assert e.spcal_amp_col[0][0] == 1 and e.spcal_amp_col[0][1] == engine.BJ_COL_CONST and e.spcal_amp_val[0][1] == 0.01
    assert e.spcal_phi_col[1][1] == 2 and e.weight_col[0][0] == 3 and e.weight_val[1][0] == 1.5
    assert m.ignored == []


# ---- parameter packing -------------------------------------------------------------------------------
def test_param_map_columns_constants_absent():
    names = ["mchirp", "q", "s1z", "cos_iota", "ra", "dec", "psi", "distance", "time_shift", "phi_ref", "foo"]
    const = {"t_gps": 1.0, "s1x": 0.0, "f_min": 20.0, "approx_name": "x"}
    m = engine.ParamMap(names, const)
    col = lambda n: m.c.column[engine.SLOT_INDEX[n]]
    assert col("mchirp") == 0 and col("q") == 1 and col("phi_ref") == 9
    assert col("t_gps") == engine.BJ_COL_CONST and m.c.value[engine.SLOT_INDEX["t_gps"]] == 1.0
    assert col("mtot") == engine.BJ_COL_ABSENT and col("iota") == engine.BJ_COL_ABSENT
    assert col("lambda1") == engine.BJ_COL_ABSENT
    assert m.ignored == ["foo"] and m.ndim == 11
    pm, x = engine.ParamMap.from_dict({"q": 2.0, "mchirp": 1.2, "zzz": 3})
    assert list(x) == [1.2, 2.0] and pm.names == ["mchirp", "q"]


# ---- set-up classes vs the oracle port -----------------------------------------------------------------
@pytest.mark.parametrize("ifo", ["H1", "L1", "V1"])
def test_noise_matches_port(ifo):
    fa, asd = design_asd(ifo)
    ours = Noise(fa, asd, f_min=20., f_max=4096.)
    ref = port.NoisePort(fa, asd, f_min=20., f_max=4096.)
    f = np.concatenate([np.arange(0, 8200, 0.37), [5., 9., 10., 8192., 9000.]])
    a, b = ours.interp_psd_pad(f), ref.interp_psd_pad(f)
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    assert np.array_equal(a[np.isfinite(a)], b[np.isfinite(b)])


@pytest.mark.parametrize("ifo", ["H1", "L1", "V1", "K1"])
def test_detector_geometry_matches_port(ifo):
    ours = Detector(ifo, syn.T_GPS)
    ref = port.DetectorPort(ifo, syn.T_GPS, ours.gmst_reference, ours.sday)
    np.testing.assert_allclose(ours.location, ref.location, rtol=1e-15)
    np.testing.assert_allclose(ours.response, ref.response, rtol=1e-14, atol=1e-16)
    rng = np.random.default_rng(3)
    for _ in range(20):
        ra, dec, psi = rng.uniform(0, 6.28), rng.uniform(-1.5, 1.5), rng.uniform(0, 3.14)
        t = syn.T_GPS + rng.uniform(-0.1, 0.1)
        assert abs(ours.time_delay_from_earth_center(ra, dec, t) - ref.time_delay_from_earth_center(ra, dec, t)) < 1e-15
        np.testing.assert_allclose(ours.antenna_pattern(ra, dec, psi, t), ref.antenna_pattern(ra, dec, psi, t),
                                   rtol=1e-12, atol=1e-15)


def test_series_band_mask_inclusive_and_store_measurement():
    cfg = syn.CONFIGS["C1"]
    f = syn.frequency_axis(cfg)
    d = (np.arange(len(f)) + 1j).astype(complex)
    s = Series("freq", d, srate=cfg["srate"], seglen=cfg["seglen"], f_min=cfg["f_min"], f_max=cfg["f_max"],
               t_gps=syn.T_GPS, only=True, importfreqs=f)
    band = f[s.mask]
    assert band[0] == cfg["f_min"] and band[-1] == cfg["f_max"] and len(band) == 4017
    assert s.window_factor == 1.0
    det = Detector("H1", syn.T_GPS)
    fa, asd = design_asd("H1")
    det.store_measurement(s, Noise(fa, asd, f_min=20., f_max=1024.))
    ref = port.DetectorPort("H1", syn.T_GPS, det.gmst_reference, det.sday)
    ref.store_measurement(port.SeriesPort(d, f, cfg["srate"], cfg["seglen"], cfg["f_min"], cfg["f_max"], syn.T_GPS),
                          port.NoisePort(fa, asd, f_min=20., f_max=1024.))
    assert np.array_equal(det.psd, ref.psd) and np.array_equal(det.data, ref.data)
    assert det._dd == ref._dd


def test_time_domain_series_fft_convention():
    srate, T = 256., 4.
    t = np.arange(int(srate * T)) / srate
    x = np.sin(2 * np.pi * 20 * t)
    s = Series("time", x, srate, seglen=T, f_min=10., f_max=100., t_gps=10., alpha_taper=0.1)
    assert len(s.freqs) == len(s.freq_series) == int(srate * T) // 2 + 1
    assert 0 < s.window_factor < 1
    k = np.argmax(np.abs(s.freq_series))
    assert abs(s.freqs[k] - 20.) < 1. / T + 1e-12


def test_gmst_closed_form_sanity():
    from bajes_b200.gmst import gmst_from_gps, gps_minus_utc
    assert gps_minus_utc(1187008882.) == 18
    g0 = gmst_from_gps(1187008882.)
    g1 = gmst_from_gps(1187008882. + 86164.09053083288)
    assert 0 <= g0 < 2 * np.pi
    assert abs((g1 - g0 + np.pi) % (2 * np.pi) - np.pi) < 1e-6
    # J2000.0 epoch (2000-01-01 12:00:00 UTC = GPS 630763213): GMST = 280.46061837 deg
    assert abs(gmst_from_gps(630763213.) - np.radians(280.46061837)) < 1e-9


# ---- pool adapter ---------------------------------------------------------------------------------------
class _FakePosterior:
    def __init__(self):
        self.calls = 0

    def log_like(self, x):
        return float(np.sum(x))

    def log_like_batch(self, X):
        self.calls += 1
        return np.sum(X, axis=1)

    def log_likeprior(self, x):
        return float(np.sum(x)), -1.0

    def log_likeprior_batch(self, X):
        self.calls += 1
        return np.sum(X, axis=1), -np.ones(len(X))


def test_batched_pool_contract():
    import multiprocessing.pool
    post = _FakePosterior()
    pool = BatchedPool(post, processes=8)
    assert isinstance(pool, multiprocessing.pool.Pool) and pool._processes == 8
    pts = [np.array([i, 2.0 * i]) for i in range(17)]
    assert pool.map(post.log_like, pts) == [post.log_like(p) for p in pts]      # order preserved
    assert post.calls == 1 and pool.n_points == 17
    assert pool.map(post.log_likeprior, pts) == [post.log_likeprior(p) for p in pts]
    assert pool.map(lambda v: v + 1, [1, 2, 3]) == [2, 3, 4]                    # unknown fn -> serial map
    assert pool.map(post.log_like, []) == []
    pool.close(); pool.join(); pool.terminate()
    assert isinstance(pickle.loads(pickle.dumps(pool)), BatchedPool)


def test_walker_draws_are_seeded_and_shaped():
    cfg = syn.CONFIGS["C2"]
    a, names = syn.draw_walkers(cfg, 100, 5, "narrow")
    b, _ = syn.draw_walkers(cfg, 100, 5, "narrow")
    assert np.array_equal(a, b) and a.shape == (100, 13) and names == syn.WALKER_NAMES
    w, _ = syn.draw_walkers(cfg, 100, 5, "wide")
    assert np.all(w[:, names.index("q")] >= 1.0)


# ---- relative-binning set-up (host NumPy) vs the oracle port ----------------------------------------------
def test_relbin_setup_matches_port():
    from bajes_b200.binning import compute_sdat, setup_bins
    cfg = syn.CONFIGS["C4_small"]
    f = syn.frequency_axis(cfg)
    fb = f[(f >= cfg["f_min"]) & (f <= cfg["f_max"])]
    n1, fbin1, ind1 = setup_bins(fb, cfg["f_min"], cfg["f_max"])
    n2, fbin2, ind2 = port.relbin_setup_bins(fb, cfg["f_min"], cfg["f_max"])
    assert n1 == n2 and np.array_equal(ind1, ind2) and np.array_equal(fbin1, fbin2)
    assert 40 <= n1 <= 80                       # the reference defaults chi=1, eps=0.5 give ~60 bins
    rng = np.random.default_rng(1)
    psds = [np.abs(rng.normal(size=len(fb))) + 1.0 for _ in range(2)]
    datas = [rng.normal(size=len(fb)) + 1j * rng.normal(size=len(fb)) for _ in range(2)]
    h0s = [rng.normal(size=len(fb)) + 1j * rng.normal(size=len(fb)) for _ in range(2)]
    a = compute_sdat(fb, fbin1, ind1, psds, datas, h0s)
    b = port.relbin_summary_data(fb, fbin2, ind2, psds, datas, h0s)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12 * np.max(np.abs(b)))


def test_bench_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (the CPU arm: oracle port on the host cores) runs without a GPU and prints ONE JSON
    line with the keys of the benchmark contract; with RANK != 0 it exits without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-points", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "walker*bins/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("likelihood evals/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert other.returncode == 0 and not [l for l in other.stdout.splitlines() if l.startswith("{")]


# ------------------------------------------------------------------------------------------------------
# product ROQ set-up (bajes_b200/roq.py) pinned to the reference golden vectors
# ------------------------------------------------------------------------------------------------------
def test_product_roq_setup_matches_reference_golden():
    """bajes_b200.roq.build_roq_dict (mirror of pipe/utils/roq.py:808-872 and pipe/__init__.py:661-716) against the
    psi weights, the omega weights at mid-grid and the logL golden vectors the UNMODIFIED reference produced
    (oracle/make_golden.py roq): same JenpyROQ files on disk, same data, same tc grid."""
    from conftest import load_golden, roq_product
    from test_oracle_golden import _roq_port
    from bajes_b200.roq import spline_tables
    gold = load_golden("roq_small.npz")
    cfg, net, pnet, roq = roq_product(gold)
    _, _, roq_port = _roq_port(gold)
    assert list(roq["mask_psi"]) == list(roq_port["mask_psi"]) and list(roq["mask_omega"]) == list(roq_port["mask_omega"])
    np.testing.assert_array_equal(roq["freqs_join"], roq_port["freqs_join"])
    for i, ifo in enumerate(cfg["ifos"]):
        np.testing.assert_allclose(roq[ifo]["psi_weights"], gold["psi"][i], rtol=1e-12)
        mid = roq[ifo]["omega_weights_interp"](cfg["seglen"] / 2 + 1e-3)
        np.testing.assert_allclose(mid, gold["omega_mid"][i], rtol=1e-9, atol=1e-11 * np.max(np.abs(gold["omega_mid"][i])))
        t, c, lo, hi = spline_tables(roq[ifo]["omega_weights_interp"])
        assert len(t) == c.shape[0] + 4 and c.shape[1] == len(roq["mask_omega"])
        assert lo == pytest.approx(cfg["seglen"] / 2 + gold["t_bounds"][0] - 0.030)
        assert hi == pytest.approx(cfg["seglen"] / 2 + gold["t_bounds"][1] + 0.030)
    # out of the tabulated range the reference's interpolant returns -inf (roq.py:866)
    assert np.all(np.isneginf(np.real(roq[cfg["ifos"][0]]["omega_weights_interp"](hi + 1.0))))
