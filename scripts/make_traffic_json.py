#!/usr/bin/env python
"""profiles/traffic.json from the ncu raw pages under gpurun_out/ (scripts/gpu_run_traffic.sh), stamped with the hash of the
kernel sources they were captured at (bench.py reports a figure only when the sources still hash to it)."""
import csv
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha():
    h = hashlib.sha1()
    for name in ("kernels.cuh", "walker.cuh", "roq_relbin.cuh", "roq_basis.cuh", "capi.cu"):
        h.update(open(os.path.join(ROOT, "bajes_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:12]


def dram(name):
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "r02_%s_raw.csv" % name))))
    v = {h: (u, x) for h, u, x in zip(rows[0], rows[1], rows[2])}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return int(sum(float(v[k][1]) * scale[v[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))


out = {"source_sha": sha()}
for key, name in (("C2:mufu:compressed", "tile_nodes"), ("C2:mufu:compressed:window_kernel", "window_nodes"),
                  ("C2:mufu", "tile_rot"), ("C2:mufu:window_kernel", "window"), ("C3:roq_wide", "roq_wide"),
                  ("C4:relbin", "relbin"), ("roq_basis:rb_project", "rb_project")):
    try:
        out[key] = dram(name)
    except Exception as e:            # capture missing: leave the key out
        print("skip", key, e)
out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none "
                "(profiles/r02_*_ncu.txt): fullgrid_tile_kernel on the node tables (all tables in one launch) and on the full "
                "grid (rotated detectors) at C2 x 4096 walkers, the FP64 window kernel of the same steps, roq_kernel at C3 x "
                "65536 points (wide box), relbin_kernel at C4 x 2^20 points, rb_project_kernel (ROQ basis greedy, 4096 x 16225 "
                "training matrix).  bench.py reports a figure only when the kernel sources hash to source_sha.")
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
