"""GPU diagnostic: per-detector <d|h> of the mixed modes against the all-FP64 GPU mode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bajes_b200 import GWLikelihood, synthetic as syn
from conftest import load_golden, port_network, product_network_from_port

def parts(like, X, names, const):
    dev = torch.device("cuda", 0)
    pm = like.param_map(names, const)
    xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    out = torch.empty(len(X), dtype=torch.float64, device=dev)
    pr = torch.empty((len(X), len(like.ifos), 3), dtype=torch.float64, device=dev)
    like.engine.loglike_device(xd.data_ptr(), len(X), X.shape[1], pm, out.data_ptr(), pr.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return pr.cpu().numpy(), out.cpu().numpy()

cfgname, gname = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("C2_small", "c2_small.npz")
cfg = syn.CONFIGS[cfgname]; gold = load_golden(gname)
net = port_network(cfg, gold); pnet = product_network_from_port(cfg, net)
names = [str(n) for n in gold["names"]]; const = syn.constants(cfg)
X = gold["X"][:-3]
ref = None
for mode in ("fp64", "poly32", "mufu"):
    like = GWLikelihood(*pnet, cfg["srate"], cfg["seglen"], cfg["approx"], sincos=mode)
    p, l = parts(like, X, names, const)
    z = p[..., 0] + 1j * p[..., 1]
    if ref is None:
        ref = (z, p[..., 2], l); continue
    rz = z / ref[0] - 1
    big = np.abs(ref[0]) > 50
    print("== %s: hh rel diff max %.2e" % (mode, np.max(np.abs(p[..., 2] / ref[1] - 1))))
    for i, ifo in enumerate(cfg["ifos"]):
        r = rz[:, i][big[:, i]]
        print("  %s: n=%d  mean(ratio-1) = %+.3e %+.3ei   std = %.2e   |dh| range %.0f..%.0f ; abs err rms %.2e max %.2e"
              % (ifo, big[:, i].sum(), r.real.mean(), r.imag.mean(), r.std(), np.abs(ref[0][:, i]).min(), np.abs(ref[0][:, i]).max(),
                 np.sqrt(np.mean(np.abs(z[:, i] - ref[0][:, i]) ** 2)), np.max(np.abs(z[:, i] - ref[0][:, i]))))
    d = l - ref[2]
    k = np.argsort(-np.abs(d))[:5]
    for j in k[:2]:
        print("   worst dlogL %+.3e at logL %.2f  sum|dh| %.1f  hh %.1f  err/|dh| %.2e" % (d[j], ref[2][j], np.abs(ref[0][j]).sum(), ref[1][j].sum(), abs(d[j]) / np.abs(ref[0][j]).sum()))
