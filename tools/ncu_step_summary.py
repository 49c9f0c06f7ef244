#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: cut out ONE train step (from the NCHW->NHWC input
conversion to the fused SGD kernel) and print per-kernel totals.  usage: ncu_step_summary.py launches.csv"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
    rows.append((r[ki], v * scale))
starts = [i for i, (k, _) in enumerate(rows) if "k_nchw_to_nhwc" in k]
ends = [i for i, (k, _) in enumerate(rows) if k.startswith("k_sgd")]
s = e = None
for a in starts:
    nxt = [b for b in ends if b > a]
    if nxt and not [c for c in starts if a < c < nxt[0] and rows[c][0] != rows[a][0]]:
        s, e = a, nxt[0]
        break
assert s is not None, "no complete step in the capture"
# the seed gradient also goes through k_nchw_to_nhwc inside the step: start at the first of the step
step = rows[s:e + 1]
tot = sum(t for _, t in step)
agg = collections.defaultdict(lambda: [0.0, 0])
for k, t in step:
    k = k.split("(")[0]
    agg[k][0] += t
    agg[k][1] += 1
print(f"# one train step: {len(step)} kernel launches, {tot:.2f} ms (serialised, cold-cache per-launch times)")
print(f"{'kernel':76s} {'ms':>8s} {'launches':>8s} {'share':>7s}")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:76]:76s} {t:8.3f} {n:8d} {100 * t / tot:6.1f}%")
