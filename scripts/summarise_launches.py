#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/*.md)."""
import csv
import io
import sys
from collections import OrderedDict


def main(path, step_kernels=("prologue", "fullgrid", "epilogue")):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, (n, t) in agg.items():
        print("| `%s` | %d | %.1f | %.1f | %.2f%% |" % (k, n, t, t / n, 100 * t / total))
    step = {k: v for k, v in agg.items() if any(s in k for s in step_kernels)}
    st = sum(v[1] for v in step.values())
    fused = sum(v[1] for k, v in step.items() if "fullgrid" in k)
    print("\nShare of the fused kernel (incl. its early-exit fix-up launch) inside one step "
          "(prologue + fullgrid + epilogue): **%.1f%%**" % (100 * fused / st))


if __name__ == "__main__":
    main(sys.argv[1])
