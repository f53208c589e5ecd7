#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output.
   python tools/ncu_source_top.py src.csv <kernel substring> [occurrence] [topn]"""
import csv
import io
import sys
from collections import defaultdict

path, pat = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
text = open(path).read().split('"Kernel Name",')
secs = [t for t in text[1:] if pat in t.split("\n", 1)[0]]
sec = secs[occ]
body = sec.split("\n", 1)[1]
rows = list(csv.DictReader(io.StringIO(body)))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("kernel:", sec.split("\n", 1)[0][:100], " instructions:", len(rows), " samples:", tot)
cls = defaultdict(lambda: [0, 0])
for r in rows:
    op = r["Source"].split()[0] if not r["Source"].startswith("@") else r["Source"].split()[1]
    op = op.split(".")[0]
    cls[op][0] += int(r["# Samples"] or 0)
    cls[op][1] += int(r["Instructions Executed"] or 0)
print("by opcode (samples, share, executed warp-instr):")
for k, v in sorted(cls.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {k:10s} {v[0]:8d} {100*v[0]/tot:5.1f}%  {v[1]:12d}")
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
print("top instructions:")
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:topn]:
    st = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"  {int(r['# Samples']):6d} {r['Source'][:70]:70s} {st}")
