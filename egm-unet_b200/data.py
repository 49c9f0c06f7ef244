"""Device-side input pipeline: the reference's SegmentationPresetTrain / SegmentationPresetEval + collate_fn
(train.py:17-54, transforms.py:30-110, my_dataset.py:103-133) as ONE libegm_b200 kernel per image, writing straight into the
batch tensors the model / Trainer consume.

The random draws are made on the host with the same generators in the same order as the reference (python `random` for the
resize size and the two flips, `torch.randint` for the crop origin), so a seeded run reproduces the reference's batches bit for
bit.  The host also builds Pillow's resample tables (libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc for the
antialiased bilinear resize, Geometry.c ImagingScaleAffine for the nearest-neighbour mask resize) in double precision; the
arithmetic on pixels happens on the GPU in integers.  No CPU fallback: pixels never pass through PIL / torchvision here.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .abi import call

PRECISION_BITS = 32 - 8 - 2
_TABLE_CACHE: Dict[Tuple[str, int, int], object] = {}


def resized_output_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision's F.resize(img, int): smaller edge -> size, the other edge int(size * long / short)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def _bilinear_tables(in_size: int, out_size: int):
    """Pillow precompute_coeffs (BILINEAR, support scaled by the down-sampling factor) + 22-bit fixed-point normalisation."""
    key = ("bil", in_size, out_size)
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)            # C (int) cast == truncation; operands are >= -0.5
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    ss = 1.0 / filterscale
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    v = np.abs((taps + xmin[:, None] - center[:, None] + 0.5) * ss)
    wgt = np.where((v < 1.0) & (taps < xmax[:, None]), 1.0 - v, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for t in range(ksize):                                                       # same left-to-right summation order as the C loop
        ww = ww + wgt[:, t]
    k = np.where(ww[:, None] != 0.0, wgt / np.where(ww == 0.0, 1.0, ww)[:, None], wgt)
    kk = (0.5 + k * float(1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)   # coefficients are >= 0 for the triangle filter
    out = (xmin.astype(np.int32), xmax.astype(np.int32), np.ascontiguousarray(kk), ksize)
    _TABLE_CACHE[key] = out
    return out


def _nearest_table(in_size: int, out_size: int) -> np.ndarray:
    """Pillow ImagingScaleAffine: xo starts at a/2 and is advanced by repeated addition of a = in/out (sequential rounding matters)."""
    key = ("nn", in_size, out_size)
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    a = float(in_size) / out_size
    xo = 0.0 + a * 0.5
    tab = np.empty(out_size, dtype=np.int32)
    for x in range(out_size):
        tab[x] = min(max(int(xo), 0), in_size - 1)
        xo += a
    _TABLE_CACHE[key] = tab
    return tab


class DevicePipeline:
    """get_transform(train) + DriveDataset.collate_fn of the reference, on the GPU.

        pipe = DevicePipeline(train=True)                       # base_size 565, crop 480, flips 0.5 / 0.5, ImageNet mean / std
        images, targets = pipe(list_of_uint8_HxWx3, list_of_uint8_HxW_masks)    # -> float32 [N,3,480,480], int64 [N,480,480] on the device
    """

    def __init__(self, train: bool, base_size: int = 565, crop_size: int = 480, hflip_prob: float = 0.5, vflip_prob: float = 0.5,
                 mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), device="cuda"):
        self.train, self.base_size, self.crop_size = train, base_size, crop_size
        self.hflip_prob, self.vflip_prob = hflip_prob, vflip_prob
        self.mean_std = torch.tensor(list(mean) + list(std), dtype=torch.float32)          # host tensor (read by the launcher)
        self.dev = torch.device(device)
        self._dev_tables: Dict[Tuple[str, int, int], tuple] = {}

    # ---- random draws, in the reference's order (transforms.py:37-39, 51, 62, 76; torchvision RandomCrop.get_params)
    def draw(self, h: int, w: int, torch_gen: Optional[torch.Generator] = None) -> dict:
        if not self.train:
            return dict(size=self.base_size, hflip=False, vflip=False, crop=None)
        size = random.randint(int(0.5 * self.base_size), int(1.2 * self.base_size))
        hflip = (random.random() < self.hflip_prob) if self.hflip_prob > 0 else False
        vflip = (random.random() < self.vflip_prob) if self.vflip_prob > 0 else False
        rh, rw = resized_output_size(h, w, size)
        cs = self.crop_size
        ph, pw = max(rh, cs), max(rw, cs)
        if ph == cs and pw == cs:
            top = left = 0
        else:
            top = int(torch.randint(0, ph - cs + 1, size=(1,), generator=torch_gen).item())
            left = int(torch.randint(0, pw - cs + 1, size=(1,), generator=torch_gen).item())
        return dict(size=size, hflip=bool(hflip), vflip=bool(vflip), crop=(top, left, cs))

    def _tables(self, kind: str, n_in: int, n_out: int):
        key = (kind, n_in, n_out)
        if key not in self._dev_tables:
            if kind == "bil":
                xmin, cnt, kk, ks = _bilinear_tables(n_in, n_out)
                self._dev_tables[key] = (torch.from_numpy(xmin).to(self.dev), torch.from_numpy(cnt).to(self.dev), torch.from_numpy(kk).to(self.dev), ks)
            else:
                self._dev_tables[key] = (torch.from_numpy(_nearest_table(n_in, n_out)).to(self.dev),)
            if len(self._dev_tables) > 4096:
                self._dev_tables.pop(next(iter(self._dev_tables)))
        return self._dev_tables[key]

    def __call__(self, images: Sequence, masks: Sequence, params: Optional[List[dict]] = None):
        imgs = [torch.as_tensor(np.ascontiguousarray(i) if not isinstance(i, torch.Tensor) else i).to(self.dev, torch.uint8).contiguous() for i in images]
        msks = [torch.as_tensor(np.ascontiguousarray(m) if not isinstance(m, torch.Tensor) else m).to(self.dev, torch.uint8).contiguous() for m in masks]
        if params is None:
            params = [self.draw(int(m.shape[0]), int(m.shape[1])) for m in msks]
        geo = []
        for m, p in zip(msks, params):
            h, w = int(m.shape[0]), int(m.shape[1])
            rh, rw = resized_output_size(h, w, p["size"])
            vh, vw = (p["crop"][2], p["crop"][2]) if p["crop"] is not None else (rh, rw)
            geo.append((h, w, rh, rw, vh, vw))
        oh, ow = max(g[4] for g in geo), max(g[5] for g in geo)                 # collate_fn: batch max size
        out_img = torch.empty(len(imgs), 3, oh, ow, dtype=torch.float32, device=self.dev)
        out_tgt = torch.empty(len(imgs), oh, ow, dtype=torch.int64, device=self.dev)
        for k, (im, mk, p, (h, w, rh, rw, vh, vw)) in enumerate(zip(imgs, msks, params, geo)):
            assert im.shape == (h, w, 3), "images are [H,W,3] uint8 RGB"
            hx = self._tables("bil", w, rw) if rw != w else (None, None, None, 0)
            vy = self._tables("bil", h, rh) if rh != h else (None, None, None, 0)
            nnx = self._tables("nn", w, rw)[0] if rw != w else None
            nny = self._tables("nn", h, rh)[0] if rh != h else None
            top, left = (p["crop"][0], p["crop"][1]) if p["crop"] is not None else (0, 0)
            call("input_transform", im, mk, h, w, rh, rw, hx[0], hx[1], hx[2], hx[3], vy[0], vy[1], vy[2], vy[3], nnx, nny,
                 int(p["hflip"]), int(p["vflip"]), top, left, vh, vw, oh, ow, self.mean_std, out_img[k], out_tgt[k])
        return out_img, out_tgt
