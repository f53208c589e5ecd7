#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel name.

    python tools/ncu_summary.py gpurun_out/launches.csv [--last N] [--cells C]
"""
import argparse
import csv
import re
from collections import OrderedDict, defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name[:110]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0, help="only the last N launches")
    ap.add_argument("--cells", type=float, default=0.0, help="cells per step, to print B/cell")
    ap.add_argument("--steps", type=int, default=1, help="steps covered by the selected launches")
    a = ap.parse_args()
    rows = OrderedDict()
    with open(a.csv) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    ids = sorted(rows)
    if a.last:
        ids = ids[-a.last:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in ids:
        d = rows[i]
        k = short(d["name"])
        g = agg[k]
        g[0] += 1
        g[1] += d.get("gpu__time_duration.sum", 0.0)
        g[2] += d.get("dram__bytes_read.sum", 0.0)
        g[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(g[1] for g in agg.values())
    print(f"{'kernel':110s} {'n':>4s} {'us':>10s} {'share':>6s} {'GB/s':>7s} {'rdMB':>8s} {'wrMB':>8s}" +
          ("  B/cell" if a.cells else ""))
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (g[2] + g[3]) / (g[1] * 1e-6) / 1e9 if g[1] else 0.0
        line = f"{k:110s} {g[0]:4d} {g[1]:10.1f} {100*g[1]/tot:5.1f}% {gbs:7.0f} {g[2]/1e6:8.1f} {g[3]/1e6:8.1f}"
        if a.cells:
            line += f" {(g[2]+g[3])/a.cells/a.steps:7.1f}"
        print(line)
    print(f"total {tot:.1f} us over {len(ids)} launches")


if __name__ == "__main__":
    main()
