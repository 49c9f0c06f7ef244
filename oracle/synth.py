"""Deterministic synthetic weights and inputs -- TEST INFRASTRUCTURE (see oracle/egm_oracle.py).

Weights are a pure function of (key name, shape), so the 25 MB state_dict never has to
be committed: the fixture generator (which runs the reference) and the GPU tests (which
run the CUDA path) both call `fill_state_dict` on their own model's `state_dict()`.
Inputs follow SURVEY.md s8(d): seed 1234, image ~ N(0,1), target in {0,1} with an
ignore band of 255 on the top rows (mimics my_dataset.py:119-122 collate padding).
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch


def _gen(key: str) -> torch.Generator:
    return torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)


def fill_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Return a new fp32/int64 CPU state_dict with the same keys/shapes as `sd`."""
    out = {}
    for k, v in sd.items():
        g = _gen(k)
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("running_var"):
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif k.endswith(".scale"):                      # RecursiveGatedAttention.scale (scalar)
            t = torch.full(shape, 1.1)
        elif len(shape) == 4:                           # conv weights: He-style so activations stay O(1)
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif len(shape) == 1 and shape[0] == 2 and k.endswith(".weight") and (
                ".h_cw." in k or ".w_hc." in k or ".c_hw." in k):
            t = torch.rand(shape, generator=g)          # MCAGate.weight ~ U[0,1)
        elif k.endswith(".weight"):                     # BN gamma
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                           # biases (BN beta, conv bias)
            t = 0.05 * torch.randn(shape, generator=g)
        out[k] = t.contiguous()
    return out


def make_inputs(n: int, h: int, w: int, seed: int = 1234, blobs: bool = False, ignore_rows: int = 8):
    """(image [n,3,h,w] fp32, target [n,h,w] int64 in {0,1,255})."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(n, 3, h, w, generator=g)
    if blobs:
        r = torch.randn(n, 1, h, w, generator=g)
        k = min(31, (min(h, w) // 2) * 2 - 1)
        target = (torch.nn.functional.avg_pool2d(r, k, 1, k // 2) > 0).long()[:, 0]
    else:
        target = torch.randint(0, 2, (n, h, w), generator=g)
    if ignore_rows:
        target[:, :min(ignore_rows, h // 4), :] = 255
    return image, target
