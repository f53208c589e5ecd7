#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python tools/show_bench.py gpurun_out/bench_*.log"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        print(f, "-- no JSON line --")
        print(open(f).read()[-1500:])
        continue
    print(f"{f}: {d['value']:.3f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.3f}  "
          f"step-frac {d.get('step_roofline', {}).get('frac', 0):.3f}  clocks {d.get('clocks')}")
    for k, v in d.get("kernels", {}).items():
        share = f"{v['share'] * 100:5.1f}%" if v.get("share") is not None else "  ovl "
        print(f"  {k:42s} {v['ms_per_step']:8.3f} ms {share} {v.get('achieved_gbs') or 0:8.0f} GB/s")
    if d.get("cpu_baseline"):
        print("  cpu_baseline", d["cpu_baseline"])
