#!/usr/bin/env python
"""Summarise a bench.py --profile-json file (EGM_PROFILE_DETAIL=1): per entry point and the slowest conv shapes."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
tot = sum(v["ms"] for v in d["kernels"].values())
print("ms/step", round(d["ms_per_step"], 2), "sum of kernels", round(tot, 2), "launches", sum(v["calls"] for v in d["kernels"].values()))
agg = collections.defaultdict(lambda: [0.0, 0])
for k, v in d["kernels"].items():
    a = agg[k.split(":")[0]]
    a[0] += v["ms"]; a[1] += v["calls"]
for n, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{n:26s} {a[0]:7.2f} ms {a[1]:4d}")
rows = []
for k, v in d["kernels"].items():
    if k.startswith("conv2d_tc") or k.startswith("conv2d_wgrad_tc"):
        a = [int(x) for x in k.split(":")[1].split(",")]
        n, h, w, ci, co, kh, kw, dil = a[-8:]          # the *_view entry points put their stride / offset ints first
        rows.append((v["ms"], k.split(":")[0][7:].replace("_view", ""), h, ci, co, kh, dil, v["calls"], 2.0 * n * h * w * ci * co * kh * kw * v["calls"] / v["ms"] / 1e9))
rows.sort(reverse=True)
for r in rows[:top]:
    print(f"  {r[1]:9s} {r[3]:4d}->{r[4]:4d} @{r[2]:3d} k{r[5]} d{r[6]:2d} x{r[7]}  {r[0]:.3f} ms  {r[8]:7.1f} TF/s")
