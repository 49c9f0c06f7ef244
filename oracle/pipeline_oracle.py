"""CPU restatement of the reference's input pipeline between the decoded image and the model input.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path.

Follows (reference file:line):
  * my_dataset.py:103-112   DriveDataset.__getitem__ : mask = PIL('L') / 255 -> clip -> PIL mode 'F'
  * train.py:17-54          SegmentationPresetTrain / Eval, get_transform(base_size=565, crop_size=480)
  * transforms.py:30-43     RandomResize : random.randint(min,max); F.resize(image, size) [PIL bilinear, antialiased];
                                           F.resize(target, size, NEAREST)
  * transforms.py:46-67     RandomHorizontalFlip / RandomVerticalFlip : random.random() < p
  * transforms.py:10-17,70-81 RandomCrop : pad_if_smaller(fill=0) on image AND target, T.RandomCrop.get_params (torch.randint), F.crop
  * transforms.py:95-110    ToTensor (uint8 -> float32 / 255; target -> int64), Normalize ((x - mean) / std in float32)
  * my_dataset.py:118-133   collate_fn / cat_list : pad to the batch max size, images with 0, targets with 255
The arithmetic of F.resize on PIL images lives in Pillow (third party; 12.2.0 in this image, the reference pins nothing):
libImaging/Resample.c (two-pass 8-bit convolution, 22-bit fixed-point coefficients, horizontal pass first) and Geometry.c
(ImagingScaleAffine nearest neighbour).  Both are restated below from their published algorithm and pinned by
tests/test_pipeline.py, which runs the reference's own transforms.py on PIL images in this container and compares bit for bit.
"""
import math
import random

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_output_size(h: int, w: int, size: int):
    """torchvision _compute_resized_output_size for an int size: smaller edge -> size, aspect kept (int truncation)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter over the full axis.
    -> (xmin[out], count[out], coeffs[out, ksize] int32)"""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, dtype=np.int32)
    cnt_a = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws, ww = [], 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            if v < 0.0:
                v = -v
            wv = 1.0 - v if v < 1.0 else 0.0
            ws.append(wv)
            ww += wv
        for x in range(xmax):
            k = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _pass(src: np.ndarray, xmin, cnt, kk, axis: int) -> np.ndarray:
    """one 8-bit resample pass along `axis` of a [H,W,C] uint8 array"""
    src = np.moveaxis(src, axis, 0).astype(np.int64)
    out = np.empty((len(xmin),) + src.shape[1:], dtype=np.uint8)
    for xx in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(int(cnt[xx])):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[H,W,C] uint8 -> [out_h,out_w,C]: ImagingResample, horizontal pass first, then vertical (each only when the size changes)."""
    h, w = img.shape[:2]
    if out_w != w:
        img = _pass(img, *bilinear_coeffs(w, out_w), axis=1)
    if out_h != h:
        img = _pass(img, *bilinear_coeffs(h, out_h), axis=0)
    return img


def nearest_table(in_size: int, out_size: int) -> np.ndarray:
    """Pillow ImagingScaleAffine (NEAREST resize): xo = a*0.5 accumulated by repeated addition of a = in/out; index = int(xo)."""
    a = float(in_size) / out_size
    xo = 0.0 + a * 0.5
    tab = np.zeros(out_size, dtype=np.int32)
    for x in range(out_size):
        xin = int(math.floor(xo)) if xo < 0 else int(xo)
        tab[x] = min(max(xin, 0), in_size - 1)
        xo += a
    return tab


def draw_params(h: int, w: int, train: bool, base_size=565, crop_size=480, hflip_prob=0.5, vflip_prob=0.5, torch_gen=None):
    """The random draws of SegmentationPresetTrain / Eval in the reference's order (python `random`, then torch.randint for the crop)."""
    import torch
    if not train:
        return dict(size=base_size, hflip=False, vflip=False, crop=None)
    size = random.randint(int(0.5 * base_size), int(1.2 * base_size))
    hflip = random.random() < hflip_prob if hflip_prob > 0 else False
    vflip = random.random() < vflip_prob if vflip_prob > 0 else False
    rh, rw = resized_output_size(h, w, size)
    ph, pw = max(rh, crop_size), max(rw, crop_size)              # pad_if_smaller (bottom / right)
    if pw == crop_size and ph == crop_size:
        top = left = 0
    else:
        top = int(torch.randint(0, ph - crop_size + 1, size=(1,), generator=torch_gen).item())
        left = int(torch.randint(0, pw - crop_size + 1, size=(1,), generator=torch_gen).item())
    return dict(size=size, hflip=bool(hflip), vflip=bool(vflip), crop=(top, left, crop_size))


def transform(img_u8: np.ndarray, mask_u8: np.ndarray, p: dict, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """img_u8 [H,W,3] uint8 RGB, mask_u8 [H,W] uint8 (PNG 'L' values) -> (float32 [3,h,w], int64 [h,w]) like the reference's dataset item."""
    h, w = mask_u8.shape
    rh, rw = resized_output_size(h, w, p["size"])
    if (rh, rw) != (h, w):
        img = pil_resize_bilinear_u8(img_u8, rh, rw)
        tgt = mask_u8[nearest_table(h, rh)][:, nearest_table(w, rw)]
    else:
        img, tgt = img_u8, mask_u8
    tgt = np.clip(tgt.astype(np.float64) / 255, 0, 255).astype(np.float32)     # my_dataset.py:106-108 (mode 'F')
    if p["hflip"]:
        img, tgt = img[:, ::-1], tgt[:, ::-1]
    if p["vflip"]:
        img, tgt = img[::-1], tgt[::-1]
    if p["crop"] is not None:
        top, left, cs = p["crop"]
        ph, pw = max(rh, cs), max(rw, cs)
        pi = np.zeros((ph, pw, 3), dtype=np.uint8); pi[:rh, :rw] = img
        pt = np.zeros((ph, pw), dtype=np.float32); pt[:rh, :rw] = tgt
        img, tgt = pi[top:top + cs, left:left + cs], pt[top:top + cs, left:left + cs]
    x = np.ascontiguousarray(img).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    x = (x - np.asarray(mean, dtype=np.float32)[:, None, None]) / np.asarray(std, dtype=np.float32)[:, None, None]
    return x.astype(np.float32), np.ascontiguousarray(tgt).astype(np.int64)


def collate(items):
    """DriveDataset.collate_fn: pad to the batch max size (top-left aligned), images with 0, targets with 255."""
    hs = max(i[0].shape[1] for i in items); ws = max(i[0].shape[2] for i in items)
    imgs = np.zeros((len(items), 3, hs, ws), dtype=np.float32)
    tgts = np.full((len(items), hs, ws), 255, dtype=np.int64)
    for k, (x, t) in enumerate(items):
        imgs[k, :, :x.shape[1], :x.shape[2]] = x
        tgts[k, :t.shape[0], :t.shape[1]] = t
    return imgs, tgts


# (H, W, train) cases of tests/golden/pipeline.npz: up- and down-scaling, portrait / landscape, smaller than the crop, no resize
GOLDEN_CASES = [(584, 565, True), (480, 640, True), (375, 500, True), (300, 260, True), (768, 1024, True), (565, 565, False),
                (584, 565, False), (333, 700, False), (640, 480, True), (900, 400, True)]


def synth_image(h: int, w: int, seed: int):
    """uint8 RGB noise-plus-structure image and a blocky 0/255 mask with a few grey (anti-aliased) pixels, from a seed."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(yy / 17.0)[..., None] * 60 + np.cos(xx / 23.0)[..., None] * 60 + 128 + rng.normal(0, 25, (h, w, 3)))
    img = np.clip(base, 0, 255).astype(np.uint8)
    blob = rng.random((h // 16 + 2, w // 16 + 2))
    mask = (np.kron(blob, np.ones((16, 16)))[:h, :w] > 0.5).astype(np.uint8) * 255
    grey = rng.random((h, w)) < 0.01
    mask[grey] = rng.integers(1, 255, int(grey.sum()), dtype=np.uint8)
    return img, mask
