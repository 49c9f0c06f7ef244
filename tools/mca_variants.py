#!/usr/bin/env python
"""Tuning aid: time egm_mca_fwd / egm_mca_bwd of the library variants built by tools/build_mca_variants.sh on the MCALayer shapes of cfg2."""
import ctypes
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(16, 240, 240, 64), (16, 120, 120, 128), (16, 60, 60, 256), (16, 30, 30, 256)]


def main():
    dev = torch.device("cuda")
    vp = ctypes.c_void_p
    for lib in sorted(glob.glob(os.path.join(ROOT, "egm-unet_b200", "variants", "libegm_*.so"))):
        L = ctypes.CDLL(lib)
        row = [os.path.basename(lib)]
        tot_f = tot_b = 0.0
        for n, h, w, c in SHAPES:
            if not L.egm_mca_fused_supported(c):
                row.append("n/a"); continue
            x = torch.relu(torch.randn(n, h, w, c, device=dev)).to(torch.bfloat16)
            dy = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
            y, du = torch.empty_like(x), torch.empty_like(x)
            idx = torch.empty(n * h * w * c, dtype=torch.uint8, device=dev)
            al4 = lambda v: (v + 3) & ~3
            gates = torch.rand(al4(n * h) + al4(n * w) + al4(n * c), device=dev)
            st = torch.cuda.current_stream().cuda_stream
            f = lambda: L.egm_mca_fwd(vp(x.data_ptr()), vp(gates.data_ptr()), vp(y.data_ptr()), vp(idx.data_ptr()), 1, n, h, w, c, vp(st))
            b = lambda: L.egm_mca_bwd(vp(x.data_ptr()), vp(gates.data_ptr()), vp(dy.data_ptr()), vp(idx.data_ptr()), vp(du.data_ptr()), 1, n, h, w, c, vp(st))
            ts = []
            for fn in (f, b):
                for _ in range(3):
                    assert fn() == 0
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 10)
            tot_f += ts[0]; tot_b += ts[1]
            gb = n * h * w * c * 2 / 1e9
            row.append(f"{h}x{c}: fwd {ts[0]*1e3:6.0f} us ({2.5*gb/ts[0]*1e3:5.0f} GB/s)  bwd {ts[1]*1e3:6.0f} us ({3.5*gb/ts[1]*1e3:5.0f} GB/s)")
        print(" | ".join(row), f"| total fwd {tot_f:.3f} ms bwd {tot_b:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
