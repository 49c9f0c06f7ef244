#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 conv kernels on the DoubleConv layer shapes (CUDA events, L2 flushed between calls).
    python tools/conv_bench.py [--shapes 32,32,480 64,64,240 ...] [--iters 5] [--kind fwd|wgrad|both]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import egm_unet_b200  # noqa: F401,E402
from egm_unet_b200.abi import call  # noqa: E402

LAYERS = [(32, 32, 480), (64, 32, 480), (32, 64, 480), (32, 64, 240), (64, 64, 240), (128, 64, 240), (64, 128, 120), (128, 128, 120),
          (256, 128, 120), (128, 256, 60), (256, 256, 60), (512, 256, 60), (256, 256, 30)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kind", default="both")
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--dil", type=int, default=1)
    a = ap.parse_args()
    shapes = [tuple(int(v) for v in s.split(",")) for s in a.shapes] if a.shapes else LAYERS
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for cin, cout, hw in shapes:
        n, k = a.n, a.k
        x = torch.randn(n, hw, hw, cin, device=dev).to(torch.bfloat16)
        dy = torch.randn(n, hw, hw, cout, device=dev).to(torch.bfloat16)
        w = torch.randn(cout, cin, k, k, device=dev)
        wf = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev)
        call("pack_conv_weight_tc", w, wf, None, cout, cin, k, k)
        y = torch.empty(n, hw, hw, cout, dtype=torch.bfloat16, device=dev)
        dw = torch.empty(w.numel(), dtype=torch.float32, device=dev)
        fl = 2.0 * n * hw * hw * cin * cout * k * k
        byt = 2.0 * n * hw * hw * (cin + cout)
        res = []
        for kind in (["fwd", "wgrad"] if a.kind == "both" else [a.kind]):
            ts = []
            for it in range(a.iters + 1):
                call("memset_zero", flush, flush.numel())
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if kind == "fwd":
                    call("conv2d_tc", x, wf, None, y, n, hw, hw, cin, cout, k, k, a.dil)
                else:
                    call("conv2d_wgrad_tc", x, dy, dw, n, hw, hw, cin, cout, k, k, a.dil)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            res.append(f"{kind} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s {byt / ms / 1e6:7.0f} GB/s(min-bytes)")
        print(f"{cin:4d}->{cout:4d} @{hw:3d}^2 k{k}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
