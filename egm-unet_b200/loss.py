"""Fused loss: `criterion` of train_utils/train_and_eval.py:7-19 as one CUDA forward+backward (csrc/loss.cu)."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import abi
from .abi import call


class _CriterionFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, logits, target, weight, ignore_index, with_dice):
        n, c, h, w = logits.shape
        lg = logits.detach().float().contiguous()
        tg = target.contiguous()
        if tg.dtype != torch.int64:
            raise TypeError("criterion: target must be int64 (as produced by the reference's DriveDataset/collate_fn)")
        ws_bytes = abi.query("loss_workspace_bytes", n, c, h, w)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lg.device)
        out = torch.empty(8, dtype=torch.float32, device=lg.device)
        need = bool(fctx.needs_input_grad[0])
        dl = torch.empty_like(lg) if need else None
        wt = None if weight is None else weight.detach().float().contiguous()
        call("loss_fwd_bwd", lg, tg, wt, n, c, h, w, int(ignore_index), int(with_dice), 1.0, out, dl, ws, ws_bytes)
        fctx.dl = dl
        fctx.terms = out
        return out[0]

    @staticmethod
    def backward(fctx, g):
        dl = fctx.dl
        fctx.dl = None
        gs = g.detach().float().contiguous()
        call("unary", dl, None, gs, dl, abi.F32, dl.numel(), 2)      # dlogits *= dL/dloss (device scalar)
        return dl, None, None, None, None


def fused_criterion(logits: torch.Tensor, target: torch.Tensor, loss_weight: Optional[torch.Tensor] = None,
                    ignore_index: int = -100, dice: bool = True) -> torch.Tensor:
    if not logits.is_cuda:
        raise RuntimeError("egm_b200 criterion runs on CUDA only (the CPU oracle lives in oracle/)")
    return _CriterionFn.apply(logits, target, loss_weight, ignore_index, dice)


def criterion(inputs: Dict[str, torch.Tensor], target, loss_weight=None, num_classes: int = 2, dice: bool = True, ignore_index: int = -100):
    """Same signature and value as the reference `criterion` (train_utils/train_and_eval.py:7-19)."""
    losses = {}
    for name, x in inputs.items():
        assert x.shape[1] == num_classes
        losses[name] = fused_criterion(x, target, loss_weight, ignore_index, dice)
    if len(losses) == 1:
        return losses["out"]
    return losses["out"] + 0.5 * losses["aux"]


def loss_terms(logits: torch.Tensor, target: torch.Tensor, loss_weight: Optional[torch.Tensor] = None, ignore_index: int = -100):
    """Forward-only: dict of the five terms (train_utils/dice_coefficient_loss.py) from one fused launch, plus `bad_labels`:
    the number of labels that are neither a class id nor `ignore_index` (the reference raises a device assert for those; the
    kernel drops them from CE and Dice alike and counts them here)."""
    n, c, h, w = logits.shape
    lg = logits.detach().float().contiguous()
    ws_bytes = abi.query("loss_workspace_bytes", n, c, h, w)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lg.device)
    out = torch.empty(8, dtype=torch.float32, device=lg.device)
    wt = None if loss_weight is None else loss_weight.detach().float().contiguous()
    call("loss_fwd_bwd", lg, target.contiguous(), wt, n, c, h, w, int(ignore_index), 1, 1.0, out, None, ws, ws_bytes)
    return {"total": out[0], "ce": out[1], "dice": out[2], "laplace": out[3], "lap": out[4], "sobel": out[5], "bad_labels": out[6]}


class EvalMetrics:
    """Fused argmax + confusion matrix + Dice accumulation (ConfusionMatrix / DiceCoefficient of
    train_utils/distributed_utils.py:76-167) -- one launch per batch, no one-hot tensors."""

    def __init__(self, num_classes: int, ignore_index: int = 255, device="cuda"):
        self.n, self.ignore = num_classes, ignore_index
        self.mat = torch.empty(num_classes * num_classes, dtype=torch.int64, device=device)
        call("memset_zero", self.mat, self.mat.numel() * 8)
        self.dice_sum = 0.0
        self.count = 0
        self._pending = []

    def update(self, logits: torch.Tensor, target: torch.Tensor):
        n, c, h, w = logits.shape
        acc = torch.empty(n * c * 3, dtype=torch.float64, device=logits.device)
        call("memset_zero", acc, acc.numel() * 8)
        call("eval_metrics", logits.detach().float().contiguous(), target.contiguous(), n, c, h, w, self.ignore, self.mat, acc)
        self._pending.append((acc, n, c))

    def _drain(self):
        for acc, n, c in self._pending:
            a = acc.cpu().view(n, c, 3)[:, 1:]                       # foreground classes only (distributed_utils.py:143)
            inter, sets = a[..., 0], a[..., 1] + a[..., 2]
            sets = torch.where(sets == 0, 2 * inter, sets)
            self.dice_sum += float(((2 * inter + 1e-6) / (sets + 1e-6)).mean())
            self.count += 1
        self._pending = []

    @property
    def dice(self) -> float:
        self._drain()
        return self.dice_sum / max(self.count, 1)

    def confusion(self) -> torch.Tensor:
        return self.mat.view(self.n, self.n)
