#!/usr/bin/env python
"""DRAM traffic of the tcgen05 conv launches of ONE train step, from an ncu launch list with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` over tools/one_step.py and the C-ABI call log of the
same step (tools/one_step.py --call-log): the i-th k_conv_tc* / k_wgrad_tc* kernel of the capture is the i-th conv2d_tc* /
conv2d_wgrad_tc* call, which names its layer -- so the DoubleConv launches (bench.py's roofline family) can be summed exactly.

    usage: ncu_conv_traffic.py launches.csv calls.txt out.json"""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rd:
        nm = r[ki].split("(")[0]
        nm = nm[5:] if nm.startswith("void ") else nm
        d = launches.setdefault(int(r[ii]), {"name": nm.split("<")[0], "full": nm})
        v = float(r[vi].replace(",", ""))
        u = r[ui].lower()
        if "time_duration" in r[mi]:
            v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
            d["ms"] = v
        else:
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
            d["dram"] = d.get("dram", 0.0) + v
    convs = [d for d in launches.values() if d["name"].startswith(("k_conv_tc", "k_wgrad_tc"))]
    calls = [ln.strip() for ln in open(sys.argv[2]) if ln.startswith(("conv2d_tc", "conv2d_wgrad_tc"))]
    assert len(convs) == len(calls), (len(convs), len(calls))
    out = {"launches_in_step": len(launches), "step_ms_serialised": sum(d.get("ms", 0.0) for d in launches.values()),
           "step_dram_bytes": sum(d.get("dram", 0.0) for d in launches.values()),
           "conv": {"launches": len(convs), "ms": sum(d["ms"] for d in convs), "dram_bytes": sum(d["dram"] for d in convs)}}
    dc = [(d, c) for d, c in zip(convs, calls) if bench.is_doubleconv(c)]
    alg = 0.0
    for d, c in dc:
        n, h, w, ci, co, kh, kw, dil = bench.conv_shape(c)
        alg += 2.0 * n * h * w * (ci + co) + 2.0 * kh * kw * ci * co + (4.0 * kh * kw * ci * co if c.startswith("conv2d_wgrad") else 0.0)
    out["doubleconv"] = {"launches": len(dc), "ms": sum(d["ms"] for d, _ in dc), "dram_bytes": sum(d["dram"] for d, _ in dc),
                         "algorithmic_bytes": alg, "note": "algorithmic = read input + write output (bf16) + weights; wgrad adds the fp32 gradient"}
    out["doubleconv_dram_bytes_per_step"] = out["doubleconv"]["dram_bytes"]        # the key bench.py reads for roofline.traffic
    out["doubleconv_algorithmic_bytes_per_step"] = alg
    per = collections.defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    for d in launches.values():
        p = per[d["name"]]
        p["launches"] += 1; p["ms"] += d.get("ms", 0.0); p["dram_bytes"] += d.get("dram", 0.0)
    out["per_kernel"] = dict(sorted(per.items(), key=lambda kv: -kv[1]["ms"]))
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    if len(sys.argv) > 4:      # human-readable per-kernel table of the step
        tot = out["step_ms_serialised"]
        with open(sys.argv[4], "w") as f:
            f.write(f"# one train step (tools/one_step.py under ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none):\n")
            f.write(f"# {out['launches_in_step']} kernel launches, {tot:.2f} ms serialised (cold-cache per-launch times), {out['step_dram_bytes'] / 1e9:.2f} GB of DRAM traffic\n")
            f.write(f"{'kernel':40s} {'ms':>8s} {'launches':>8s} {'share':>7s} {'DRAM GB':>9s} {'GB/s':>8s}\n")
            for k, v in out["per_kernel"].items():
                f.write(f"{k[:40]:40s} {v['ms']:8.3f} {v['launches']:8d} {100 * v['ms'] / tot:6.1f}% {v['dram_bytes'] / 1e9:9.3f} {v['dram_bytes'] / max(v['ms'], 1e-9) / 1e6:8.0f}\n")
    print(json.dumps({k: out[k] for k in ("launches_in_step", "step_ms_serialised", "step_dram_bytes", "conv", "doubleconv")}, indent=1))


if __name__ == "__main__":
    main()
