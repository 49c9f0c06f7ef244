#!/usr/bin/env python
"""Measure the CLIPSeg-ensemble alpha sweep: libegm_b200 vs the CPU oracle port on the same synthetic
validation set.   python tools/ensemble_bench.py [--images 16] [--size 480 640]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import egm_unet_b200  # noqa: F401,E402
from egm_unet_b200 import ensemble as ENS  # noqa: E402
from oracle import ensemble_oracle as EO  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--size", type=int, nargs=2, default=[480, 640])
    ap.add_argument("--cpu-images", type=int, default=2, help="images of the same set timed through the CPU oracle (bounded sample)")
    a = ap.parse_args()
    h, w = a.size
    clip, unet, labels = EO.make_case(11, [((h, w), (h, w))] * a.images)
    alphas = np.linspace(0.1, 10.0, 100)
    dev = torch.device("cuda")
    clip_d = [torch.from_numpy(c)[None].to(dev) for c in clip]
    unet_d = [torch.from_numpy(u)[None].to(dev) for u in unet]
    lab_d = [torch.from_numpy(l).to(dev) for l in labels]
    for _ in range(3):
        ENS.alpha_sweep(clip_d, unet_d, lab_d, alphas)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        conf, miou, best = ENS.alpha_sweep(clip_d, unet_d, lab_d, alphas)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = a.images * (2 * 352 * 352 * 4 + 2 * h * w * 4 + h * w)
    t0 = time.perf_counter()
    EO.search_best_alpha(clip[: a.cpu_images], unet[: a.cpu_images], labels[: a.cpu_images])
    cpu_ms = (time.perf_counter() - t0) * 1e3 / a.cpu_images * a.images
    print(f"alpha sweep, {a.images} images {h}x{w}, 100 alphas: GPU {ms:.3f} ms ({a.images / ms * 1e3:.0f} images/s, {byt / ms / 1e6:.1f} GB/s of "
          f"algorithmic bytes, {a.images * h * w * 100 / ms / 1e6:.1f} G pixel-alphas/s); CPU oracle port (numpy, 1 core, extrapolated from "
          f"{a.cpu_images} images) {cpu_ms:.0f} ms; best alpha {float(best[0]):.4f} mIoU {float(best[1]):.4f}")


if __name__ == "__main__":
    main()
