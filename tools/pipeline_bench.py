#!/usr/bin/env python
"""Measure the device input pipeline: one training batch of 16 decoded 584x565 uint8 images -> normalised
[16,3,480,480] float32 + int64 targets.  GPU (libegm_b200, CUDA events, images already on the device and, separately, including the
H2D copy of the uint8 images) vs the reference's library path (PIL + torchvision functional ops, one host core per worker)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import egm_unet_b200  # noqa: F401,E402
from egm_unet_b200.data import DevicePipeline  # noqa: E402
from oracle import pipeline_oracle as PO  # noqa: E402


def pil_path(img, mask, p, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """what transforms.py executes per sample (torchvision functional calls on PIL images), with the draws fixed to p"""
    from PIL import Image
    from torchvision import transforms as T
    from torchvision.transforms import functional as F
    im = Image.fromarray(img)
    tg = Image.fromarray(np.clip(np.array(Image.fromarray(mask)) / 255, 0, 255))
    im = F.resize(im, p["size"]); tg = F.resize(tg, p["size"], interpolation=T.InterpolationMode.NEAREST)
    if p["hflip"]:
        im, tg = F.hflip(im), F.hflip(tg)
    if p["vflip"]:
        im, tg = F.vflip(im), F.vflip(tg)
    cs = p["crop"][2]
    ow, oh = im.size
    if min(ow, oh) < cs:
        pad = (0, 0, max(cs - ow, 0), max(cs - oh, 0))
        im, tg = F.pad(im, pad, fill=0), F.pad(tg, pad, fill=0)
    im, tg = F.crop(im, p["crop"][0], p["crop"][1], cs, cs), F.crop(tg, p["crop"][0], p["crop"][1], cs, cs)
    x = F.normalize(F.to_tensor(im), mean=mean, std=std)
    return x, torch.as_tensor(np.array(tg), dtype=torch.int64)


def main():
    n, h, w = 16, 584, 565
    data = [PO.synth_image(h, w, 300 + i) for i in range(n)]
    pipe = DevicePipeline(train=True)
    random.seed(1); torch.manual_seed(1)
    params = [pipe.draw(h, w) for _ in range(n)]
    dev = torch.device("cuda")
    imgs_d = [torch.from_numpy(d[0]).to(dev) for d in data]
    msks_d = [torch.from_numpy(d[1]).to(dev) for d in data]
    imgs_h = [torch.from_numpy(d[0]).pin_memory() for d in data]
    msks_h = [torch.from_numpy(d[1]).pin_memory() for d in data]
    for _ in range(3):
        pipe(imgs_d, msks_d, params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        x, t = pipe(imgs_d, msks_d, params)
    e1.record(); torch.cuda.synchronize()
    ms_res = e0.elapsed_time(e1) / reps
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        x, t = pipe([i.to(dev, non_blocking=True) for i in imgs_h], [m.to(dev, non_blocking=True) for m in msks_h], params)
    e1.record(); torch.cuda.synchronize()
    ms_h2d = e0.elapsed_time(e1) / reps
    pil_path(data[0][0], data[0][1], params[0])          # warm-up (imports, PIL plugin init)
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    ref = [pil_path(d[0], d[1], p) for d, p in zip(data[:8], params[:8])]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 8 * n
    same = all(torch.equal(x[k].cpu(), ref[k][0]) and torch.equal(t[k].cpu(), ref[k][1]) for k in range(4))
    out_bytes = n * (3 * 480 * 480 * 4 + 480 * 480 * 8); in_bytes = n * (h * w * 4)
    print(f"input pipeline, batch {n} of {h}x{w} uint8 -> [16,3,480,480] f32 + int64 targets: GPU {ms_res:.3f} ms resident "
          f"({n / ms_res * 1e3:.0f} images/s, {(in_bytes + out_bytes) / ms_res / 1e6:.1f} GB/s of in+out bytes), {ms_h2d:.3f} ms incl. H2D of the uint8 images; "
          f"PIL + torchvision on 1 host core {cpu_ms:.1f} ms ({n / cpu_ms * 1e3:.0f} images/s); outputs identical to the PIL path: {same}")


if __name__ == "__main__":
    main()
