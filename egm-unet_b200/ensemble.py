"""CLIPSeg-ensemble fusion step on the GPU: the stage immediately downstream of the UNet logits in the
reference's final pipeline.  Host-side mirror of the reference functions, same names and argument meaning:

  * search_best_alpha(clip_logits_list, unet_logits_list, labels_list, search_scale=[0.1, 10.0], search_step=100)
        -- eval_CLIPseg.py:656-724.  The CLIP logits may be the raw [1,C,352,352] CLIPSeg output or already interpolated to the UNet
           size (eval_CLIPseg.py:885-888): the bilinear resize is fused into the kernel and is the identity for equal sizes.
  * fuse_predict(clip_logits, unet_logits, alpha, original_size)
        -- eval_CLIPseg.py:901-912 / predict_CLIPseg.py:519-526: uint8 mask at the original image size (PIL size = (width, height)).

All arithmetic runs in libegm_b200 (csrc/ensemble.cu): one launch per image covers every alpha.  No CPU fallback.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .abi import call


def _plane(t, dev) -> torch.Tensor:
    """[1,C,H,W] / [C,H,W] float tensor or array -> contiguous fp32 [C,H,W] on the device."""
    t = torch.as_tensor(t)
    if t.dim() == 4:
        assert t.shape[0] == 1, "one image per list entry (as in the reference)"
        t = t[0]
    return t.to(device=dev, dtype=torch.float32).contiguous()


def alpha_sweep(clip_logits_list: Sequence, unet_logits_list: Sequence, labels_list: Sequence, alphas: np.ndarray, num_classes: int = 2,
                device="cuda") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (confusion [n_alpha,C,C] int64, mIoU per alpha float32, best {alpha, mIoU} float64), all on the device."""
    dev = torch.device(device)
    a_dev = torch.as_tensor(np.asarray(alphas, dtype=np.float64)).to(dev)
    n_alpha = a_dev.numel()
    conf = torch.zeros(n_alpha * num_classes * num_classes, dtype=torch.int64, device=dev)
    keep = []
    for clip, unet, lab in zip(clip_logits_list, unet_logits_list, labels_list):
        c, u = _plane(clip, dev), _plane(unet, dev)
        lab_t = lab if isinstance(lab, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(lab))
        lab_t = lab_t.to(device=dev, dtype=torch.uint8).contiguous()
        assert c.shape[0] == num_classes and u.shape[0] == num_classes and lab_t.dim() == 2
        call("ensemble_confusion", c, c.shape[1], c.shape[2], u, u.shape[1], u.shape[2], lab_t, lab_t.shape[0], lab_t.shape[1], num_classes,
             a_dev, n_alpha, conf)
        keep.append((c, u, lab_t))                 # keep inputs alive until the stream has consumed them
    miou = torch.empty(n_alpha, dtype=torch.float32, device=dev)
    best = torch.empty(2, dtype=torch.float64, device=dev)
    call("ensemble_best_alpha", conf, a_dev, n_alpha, num_classes, miou, best)
    torch.cuda.current_stream(dev).synchronize()
    return conf.view(n_alpha, num_classes, num_classes), miou, best


def search_best_alpha(clip_logits_list: List, unet_logits_list: List, labels_list: List, search_scale=[0.1, 10.0], search_step=100) -> float:
    alpha_min, alpha_max = search_scale
    alphas = np.linspace(alpha_min, alpha_max, search_step)
    _, _, best = alpha_sweep(clip_logits_list, unet_logits_list, labels_list, alphas)
    return float(best[0])


def fuse_predict(clip_logits, unet_logits, alpha: float, original_size: Tuple[int, int], device="cuda") -> np.ndarray:
    """uint8 mask [height, width]; original_size follows PIL's (width, height) convention like the reference's cv2.resize call."""
    dev = torch.device(device)
    c, u = _plane(clip_logits, dev), _plane(unet_logits, dev)
    wo, ho = int(original_size[0]), int(original_size[1])
    mask = torch.empty(ho, wo, dtype=torch.uint8, device=dev)
    call("ensemble_predict", c, c.shape[1], c.shape[2], u, u.shape[1], u.shape[2], c.shape[0], float(np.float32(alpha)), mask, ho, wo)
    return mask.cpu().numpy()
