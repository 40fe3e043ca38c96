import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "needs_reference: needs the read-only reference tree (build container only)")


def _have_gpu():
    try:
        from bajes_b200 import engine
        engine.device_info(0)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    from oracle import shims
    have_ref = shims.reference_available()
    skip_ref = pytest.mark.skip(reason="reference tree not present on this machine")
    gpu_items = [item for item in items if "gpu" in item.keywords]
    skip_gpu = None
    if gpu_items and not _have_gpu():
        # a plain `pytest` on a CPU box skips the GPU tests; the product itself still fails loudly without a device
        skip_gpu = pytest.mark.skip(reason="no B200 / CUDA library on this machine (bajes_b200 has no CPU fallback)")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if skip_gpu is not None and "gpu" in item.keywords:
            item.add_marker(skip_gpu)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure the in-tree CUDA library exists (compiles in seconds, no GPU needed)."""
    import __graft_entry__ as g
    if g._stale():
        g.build()
    yield


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def port_network(cfg, gold, with_golden_data=False):
    """Re-synthesise the injection with the oracle port classes, pinned to the detector constants stored in
    the golden file (gmst_reference / sday as produced when the reference ran)."""
    from bajes_b200 import synthetic as syn
    from oracle import harness
    cl = harness.port_classes(gmst_reference=float(gold["gmst_reference"][0]), sday=float(gold["sday"][0]))
    return syn.build_network(cfg, cl)


def product_network_from_port(cfg, net):
    """Product (GPU) objects carrying EXACTLY the data / detector constants of a port network, so both sides
    of a parity test see identical inputs."""
    from bajes_b200 import synthetic as syn
    cl = syn.product_classes()
    ifos, datas_p, dets_p, noises_p, f = net
    datas, dets, noises = {}, {}, {}
    for ifo in ifos:
        d = cl.Detector(ifo, syn.T_GPS)
        d.gmst_reference, d.sday = dets_p[ifo].gmst_reference, dets_p[ifo].sday
        dets[ifo] = d
        datas[ifo] = cl.make_series(datas_p[ifo].freq_series, f, cfg)
        fa, asd = design_asd_any(ifo)
        noises[ifo] = cl.Noise(fa, asd, f_min=cfg["f_min"], f_max=cfg["f_max"])
    return ifos, datas, dets, noises, f


def design_asd_any(ifo):
    """design curve of the detector, the aLIGO one for sites that have no bundled table (tests of larger networks)"""
    from bajes_b200.noise import design_asd
    try:
        return design_asd(ifo)
    except AttributeError:
        return design_asd("H1")


def roq_product(gold, device=None):
    """(cfg, port network, product network, roq dict) for the reduced ROQ case of tests/golden/roq_small.npz, the roq
    dict assembled by the PRODUCT set-up (bajes_b200.roq.build_roq_dict) from a JenpyROQ directory on disk."""
    import tempfile
    from bajes_b200 import roq as roqmod, synthetic as syn
    from bajes_b200.ensemble import BoxPrior
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=float(gold["seglen"]), seed=int(gold["seed"]))
    net = port_network(cfg, gold)
    pnet = product_network_from_port(cfg, net)
    ifos, datas, dets, noises, f = pnet
    names = list(syn.WALKER_NAMES)
    bounds = [[-1.0, 1.0]] * len(names)
    prior = BoxPrior(names, bounds, {})
    prior.bounds[names.index("time_shift")] = [float(gold["t_bounds"][0]), float(gold["t_bounds"][1])]
    basis = syn.make_roq_basis(cfg, n_lin=48, n_quad=24, seed=7)
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, basis, cfg)
        roq = roqmod.build_roq_dict(tmp, int(gold["time_points"]), ifos, datas, noises, f, cfg["f_min"], cfg["f_max"],
                                    cfg["seglen"], prior, cfg["approx"], device=device)
    return cfg, net, pnet, roq
