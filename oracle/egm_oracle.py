"""CPU oracle for the EGM-UNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (torch CPU ops, autograd for the backward) of the
reference's live graph: `UNet` (src/unet.py:61-96), the EGM-UNet `GRFBUNet`
(src/EGM-UNet.py:1503-1541), its `yuanGRFBUNet` variant
(src/yuanGRFBUNet.py:859-875: DoubleConv1 without MCALayer) and the loss
`criterion` (train_utils/train_and_eval.py:7-19, train_utils/dice_coefficient_loss.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file.  It never touches CUDA.

Parity pinning: the reference ships no tests or golden vectors, so
this restatement is pinned against the reference *source executed in the
authoring container* (`oracle/gen_golden.py` imports /root/reference by path,
asserts this file reproduces it to fp32 round-off on logits, loss and every
parameter gradient, and writes the fixtures in tests/golden/).

All functions take a flat `state_dict`-style mapping `sd` with the reference's
exact key names (e.g. `down1.1.7.branch_edge.2.conv.weight`) and NCHW fp32
tensors.  `bn_updates`, when given, collects the running-stat updates
(training mode) so tests can compare buffers after a step.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #
# Optional storage-rounding model: with STORAGE = torch.bfloat16 every activation tensor the CUDA path materialises
# (conv outputs, BN+activation outputs, block outputs) is rounded to bf16 and back, arithmetic staying fp32.  This is the
# "reference computed with bf16-stored activations" the bf16 parity tests compare against (tests/test_model_gpu.py).
STORAGE = None
# error-budget switch (tools/bf16_error_budget.py): storage classes listed here stay fp32 even when STORAGE is set.
# classes: "dc_z" DoubleConv conv outputs (pre-BN), "dc_y" DoubleConv BN+ReLU outputs, "bc_z"/"bc_y" the same for the GRFB BasicConvs,
# "conv" other conv outputs (FusionConv, shortcut, RGA, out_conv inputs), "edge" edge-enhancer tensors, "mca" MCALayer output,
# "mix" FusionConv f+s*ca, "grfb" GRFB residual / gated outputs, "rga" RGA gate tensors, "up" upsampled tensors, "input",
# "weight" conv weights as MMA operands
STORAGE_FP32 = frozenset()
STORAGE_FP16 = frozenset()      # classes stored as IEEE fp16 (11-bit significand) instead of STORAGE


def _r(x: Tensor, tag: str = "conv") -> Tensor:
    if STORAGE is None or tag in STORAGE_FP32:
        return x
    if tag in STORAGE_FP16:
        return x.to(torch.float16).to(x.dtype)
    return x.to(STORAGE).to(x.dtype)


def _bn(sd, p: str, x: Tensor, train: bool, momentum: float, upd: Optional[dict], eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d: batch stats + biased var for normalisation in
    training, running stats in eval; unbiased var into running_var."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    if train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if upd is not None:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                upd[p + ".running_mean"] = (1 - momentum) * sd[p + ".running_mean"] + momentum * mean
                upd[p + ".running_var"] = (1 - momentum) * sd[p + ".running_var"] + momentum * var * n / max(n - 1, 1)
                upd[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    scale = w / torch.sqrt(var + eps)
    return x * scale[None, :, None, None] + (b - mean * scale)[None, :, None, None]


def _conv(sd, p: str, x: Tensor, padding=0, dilation=1, groups=1, tag: str = "conv") -> Tensor:
    # the tcgen05 path feeds the conv weights to the MMA as bf16 operands: the storage model rounds them too (class "weight")
    return _r(F.conv2d(x, _r(sd[p + ".weight"], "weight"), sd.get(p + ".bias"), 1, padding, dilation, groups), tag)


def double_conv(sd, p, x, train, upd, i0=0, i1=3):
    """DoubleConv: src/unet.py:7-18 == src/EGM-UNet.py:44-55 (conv idx i0,i1; BN idx +1)."""
    x = _r(F.relu(_bn(sd, f"{p}.{i0 + 1}", _conv(sd, f"{p}.{i0}", x, 1, tag="dc_z"), train, 0.1, upd)), "dc_y")
    x = _r(F.relu(_bn(sd, f"{p}.{i1 + 1}", _conv(sd, f"{p}.{i1}", x, 1, tag="dc_z"), train, 0.1, upd)), "dc_y")
    return x


def basic_conv(sd, p, x, train, upd, padding=0, dilation=1, groups=1, relu=True):
    """BasicConv: src/EGM-UNet.py:958-975 (BN momentum 0.01)."""
    x = _bn(sd, p + ".bn", _conv(sd, p + ".conv", x, padding, dilation, groups, tag="bc_z"), train, 0.01, upd)
    return _r(F.relu(x) if relu else x, "bc_y")


def edge_enhancer(sd, p, x, train, upd):
    """EdgeAwareFeatureEnhancer: src/EGM-UNet.py:872-886."""
    e = _r(x - F.avg_pool2d(x, 3, 1, 1), "edge")          # count_include_pad=True -> /9
    z = _conv(sd, p + ".weight_generator.0", e, tag="edge")
    w = torch.sigmoid(_bn(sd, p + ".weight_generator.1", z, train, 0.1, upd))
    return _r(w * x + x, "edge")


def mca_gate(sd, p, x):
    """MCAGate: src/EGM-UNet.py:836-869. x [B,C',H',W'] -> gate over C' from (avg,std) over H'W'."""
    b, c = x.shape[:2]
    avg = x.mean(dim=(2, 3), keepdim=True)
    std = x.reshape(b, c, -1).std(dim=2, keepdim=True).reshape(b, c, 1, 1)   # unbiased
    wt = torch.sigmoid(sd[p + ".weight"])
    out = 0.5 * (avg + std) + wt[0] * avg + wt[1] * std
    out = out.permute(0, 3, 2, 1)                                             # [B,1,1,C']
    k = sd[p + ".conv.weight"].shape[-1]
    out = F.conv2d(out, sd[p + ".conv.weight"], None, 1, (0, (k - 1) // 2))
    out = torch.sigmoid(out.permute(0, 3, 2, 1))
    return x * out


def mca_layer(sd, p, x):
    """MCALayer: src/EGM-UNet.py:686-791. frequency_enhancement == 1.1*x."""
    x_h = mca_gate(sd, p + ".h_cw", x.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    x_w = mca_gate(sd, p + ".w_hc", x.permute(0, 3, 2, 1)).permute(0, 3, 2, 1)
    x_c = mca_gate(sd, p + ".c_hw", x)
    u = (1.0 / 3.0) * (x_c + x_h + x_w)
    rng = F.max_pool2d(u, 3, 1, 1) + F.max_pool2d(-u, 3, 1, 1)               # max - min
    mean = F.avg_pool2d(u, 3, 1, 1)
    var = F.avg_pool2d((u - mean) ** 2, 3, 1, 1)
    n, c, h, w = u.shape
    shuf = u.view(n, 4, c // 4, h, w).transpose(1, 2).reshape(n, c, h, w)
    return _r(0.4 * u + 0.2 * rng + 0.2 * var + 0.1 * (1.1 * u) + 0.1 * shuf, "mca")


def fusion_conv(sd, p, x):
    """FusionConv (x1 is x2): src/EGM-UNet.py:1202-1236 with the attention modules :1171-1200."""
    f = _conv(sd, p + ".down", torch.cat([x, x], 1))
    s = _conv(sd, p + ".conv_3x3", f, 1) + _conv(sd, p + ".conv_5x5", f, 2) + _conv(sd, p + ".conv_7x7", f, 3)
    mm = torch.cat([s.mean(1, keepdim=True), s.max(1, keepdim=True)[0]], 1)
    s = s * torch.sigmoid(F.conv2d(mm, sd[p + ".spatial_attention.conv1.weight"], None, 1, 3))
    def mlp(v):
        return F.conv2d(F.relu(F.conv2d(v, sd[p + ".channel_attention.fc.0.weight"])), sd[p + ".channel_attention.fc.2.weight"])
    ca = torch.sigmoid(mlp(F.adaptive_avg_pool2d(f, 1)) + mlp(F.adaptive_max_pool2d(f, 1)))
    return _conv(sd, p + ".up", _r(f + s * ca, "mix"))


def grfb(sd, p, x, train, upd, visual=12, scale=0.1):
    """EdgeEnhancedGRFB: src/EGM-UNet.py:1238-1323."""
    xe = edge_enhancer(sd, p + ".edge_enhancer", x, train, upd)
    inter = sd[p + ".branch_edge.0.conv.weight"].shape[0]
    d = basic_conv(sd, p + ".branch_dir.0", xe, train, upd)
    d = basic_conv(sd, p + ".branch_dir.1", d, train, upd, visual, visual, relu=False)
    d = basic_conv(sd, p + ".branch_dir.2", d, train, upd)
    e = basic_conv(sd, p + ".branch_edge.0", xe, train, upd)
    e = edge_enhancer(sd, p + ".branch_edge.1", e, train, upd)
    e = basic_conv(sd, p + ".branch_edge.2", e, train, upd, 1, 1, inter)
    e = basic_conv(sd, p + ".branch_edge.3", e, train, upd, 2 * visual, 2 * visual, relu=False)
    e = basic_conv(sd, p + ".branch_edge.4", e, train, upd)
    c = basic_conv(sd, p + ".branch_ctx.0", xe, train, upd, 1)
    c = basic_conv(sd, p + ".branch_ctx.1", c, train, upd, 1, 1, 2)
    c = basic_conv(sd, p + ".branch_ctx.2", c, train, upd, 3 * visual, 3 * visual, relu=False)
    c = basic_conv(sd, p + ".branch_ctx.3", c, train, upd)
    cat = torch.cat([x, d, e, c], 1)
    out = fusion_conv(sd, p + ".fusion_conv", cat)
    zs = _conv(sd, p + ".shortcut.conv", x)
    out = _r(F.relu(out * scale + _bn(sd, p + ".shortcut.bn", zs, train, 0.01, upd)), "grfb")
    t = torch.sigmoid(_conv(sd, p + ".target_enhancer.0", out, 1))
    return _r(out * (1 + t.mean(1, keepdim=True)), "grfb")


def rga(sd, p, x):
    """RecursiveGatedAttention(order=2): src/EGM-UNet.py:458-547."""
    dim = x.shape[1]
    half = dim // 2
    fused = _conv(sd, p + ".proj_in", x)
    base, gates = fused[:, :half], fused[:, half:]
    gates = _r(_conv(sd, p + ".dwconv", gates, 1, 1, gates.shape[1]) * sd[p + ".scale"], "rga")
    out = base
    for i in range(2):
        g = gates[:, i * half:(i + 1) * half]
        g = _r(F.gelu(_conv(sd, f"{p}.gate_convs.{i}.0", g)), "rga")
        g = torch.sigmoid(_conv(sd, f"{p}.gate_convs.{i}.2", g))
        out = _r(out * g, "rga")
        if i == 0:
            out = _conv(sd, p + ".transform_convs.0", out)
    return _conv(sd, p + ".proj_out", out)


def up_block(sd, p, x1, x2, train, upd):
    """Up (bilinear): src/unet.py:29-51 == src/EGM-UNet.py:927-949."""
    if (p + ".up.weight") in sd:                                     # bilinear=False (UNet only)
        x1 = F.conv_transpose2d(x1, sd[p + ".up.weight"], sd[p + ".up.bias"], 2)
    else:
        x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = _r(F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2]), "up")
    return double_conv(sd, p + ".conv", torch.cat([x2, x1], 1), train, upd)


def down_block(sd, p, x, train, upd, variant):
    """Down: src/unet.py:21-26 (variant 'unet'), src/EGM-UNet.py:888-912 ('egm'),
    src/yuanGRFBUNet.py:859-875 ('yuan')."""
    x = F.max_pool2d(x, 2, 2)
    q = p + ".1"
    if variant == "unet":
        return double_conv(sd, q, x, train, upd)
    x = _r(F.relu(_bn(sd, q + ".1", _conv(sd, q + ".0", x, 1, tag="dc_z"), train, 0.1, upd)), "dc_y")
    if variant == "egm":
        x = mca_layer(sd, q + ".3", x)
        c2, g = 4, 7
    else:
        c2, g = 3, 6
    x = _r(F.relu(_bn(sd, f"{q}.{c2 + 1}", _conv(sd, f"{q}.{c2}", x, 1, tag="dc_z"), train, 0.1, upd)), "dc_y")
    return grfb(sd, f"{q}.{g}", x, train, upd)


def forward(sd: Dict[str, Tensor], x: Tensor, variant: str = "egm", train: bool = True,
            bn_updates: Optional[dict] = None) -> Tensor:
    """Whole-model forward -> logits [N,num_classes,H,W]. variant in {'unet','egm','yuan'}."""
    x1 = double_conv(sd, "in_conv", _r(x, "input"), train, bn_updates)
    x2 = down_block(sd, "down1", x1, train, bn_updates, variant)
    x3 = down_block(sd, "down2", x2, train, bn_updates, variant)
    x4 = down_block(sd, "down3", x3, train, bn_updates, variant)
    x5 = down_block(sd, "down4", x4, train, bn_updates, variant)
    if variant != "unet":
        x5 = rga(sd, "attn1", x5)
    y = up_block(sd, "up1", x5, x4, train, bn_updates)
    y = up_block(sd, "up2", y, x3, train, bn_updates)
    y = up_block(sd, "up3", y, x2, train, bn_updates)
    y = up_block(sd, "up4", y, x1, train, bn_updates)
    # OutConv: the CUDA path computes the logits from the (stored) activation with fp32 weights and writes them in fp32 (csrc/outconv.cu),
    # so the storage model rounds neither the weight nor the result here
    return F.conv2d(y, sd["out_conv.0.weight"], sd.get("out_conv.0.bias"))


# --------------------------------------------------------------------------- #
# loss  (train_utils/train_and_eval.py:7-19 + dice_coefficient_loss.py)
# --------------------------------------------------------------------------- #
_LAP4 = torch.tensor([[0., 1, 0], [1, -4, 1], [0, 1, 0]]).view(1, 1, 3, 3)
_LAP8 = torch.tensor([[-1., -1, -1], [-1, 8, -1], [-1, -1, -1]]).view(1, 1, 3, 3)
_SOBX = torch.tensor([[1., 0, -1], [2, 0, -2], [1, 0, -1]]).view(1, 1, 3, 3)
_SOBY = torch.tensor([[1., 2, 1], [0, 0, 0], [-1, -2, -1]]).view(1, 1, 3, 3)


def loss_terms(logits: Tensor, target: Tensor, loss_weight=None, num_classes: int = 2,
               ignore_index: int = 255) -> Dict[str, Tensor]:
    """The five terms of `criterion` (dice=True).  Vectorised, but term-by-term identical to
    dice_coefficient_loss.py:22-108 including its quirks: lap/sobel use only sample 0's
    target (raw 255s included) broadcast over the batch; Dice averages per (sample, class)
    over non-ignored pixels with the `sets_sum == 0 -> 2*inter` substitution."""
    n, c = logits.shape[:2]
    ce = F.cross_entropy(logits, target, ignore_index=ignore_index, weight=loss_weight)
    p = F.softmax(logits, dim=1)
    valid = (target != ignore_index)
    tt = torch.where(valid, target, torch.zeros_like(target))
    onehot = F.one_hot(tt, num_classes).permute(0, 3, 1, 2).to(p.dtype)
    vm = valid[:, None].to(p.dtype)
    inter = (p * onehot * vm).flatten(2).sum(2)                      # [N,C]
    sets = (p * vm).flatten(2).sum(2) + (onehot * vm).flatten(2).sum(2)
    sets = torch.where(sets == 0, 2 * inter, sets)
    dice = ((2 * inter + 1e-6) / (sets + 1e-6)).mean(0).mean(0)
    x0 = logits[:, 0:1]
    dt = logits.dtype
    t0 = target.to(dt)[0:1, None]
    k4, k8, kx, ky = _LAP4.to(dt), _LAP8.to(dt), _SOBX.to(dt), _SOBY.to(dt)
    lap4 = F.conv2d(x0, k4, padding=1).abs().mean()
    lap8 = (F.conv2d(x0, k8, padding=1) - F.conv2d(t0, k8, padding=1)).abs().mean()
    sob = ((F.conv2d(x0, kx, padding=1) - F.conv2d(t0, kx, padding=1)).abs()
           + (F.conv2d(x0, ky, padding=1) - F.conv2d(t0, ky, padding=1)).abs()).mean()
    return {"ce": ce, "dice": 1 - dice, "laplace": lap4, "lap": lap8, "sobel": sob}


def criterion(logits: Tensor, target: Tensor, loss_weight=None, num_classes: int = 2,
              ignore_index: int = 255) -> Tensor:
    t = loss_terms(logits, target, loss_weight, num_classes, ignore_index)
    return t["ce"] + t["dice"] + t["laplace"] + t["lap"] + t["sobel"]


# --------------------------------------------------------------------------- #
# eval metrics (train_utils/distributed_utils.py:76-167)
# --------------------------------------------------------------------------- #
def confusion_matrix(target: Tensor, pred: Tensor, n: int) -> Tensor:
    a, b = target.flatten(), pred.flatten()
    k = (a >= 0) & (a < n)
    return torch.bincount(n * a[k].to(torch.int64) + b[k], minlength=n * n).reshape(n, n)


def miou(mat: Tensor) -> float:
    h = mat.float()
    iu = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
    return iu.mean().item()


def dice_metric(logits: Tensor, target: Tensor, num_classes: int = 2, ignore_index: int = 255) -> float:
    """DiceCoefficient.update for one batch: one-hot argmax vs target, foreground classes."""
    pred = F.one_hot(logits.argmax(1), num_classes).permute(0, 3, 1, 2).float()
    valid = (target != ignore_index)
    tt = torch.where(valid, target, torch.zeros_like(target))
    onehot = F.one_hot(tt, num_classes).permute(0, 3, 1, 2).float()
    vm = valid[:, None].float()
    inter = (pred * onehot * vm).flatten(2).sum(2)[:, 1:]
    sets = ((pred * vm).flatten(2).sum(2) + (onehot * vm).flatten(2).sum(2))[:, 1:]
    sets = torch.where(sets == 0, 2 * inter, sets)
    return ((2 * inter + 1e-6) / (sets + 1e-6)).mean().item()


# --------------------------------------------------------------------------- #
# one SGD train step (train.py:113-118 semantics), used by the CPU baseline
# --------------------------------------------------------------------------- #
def train_step(sd: Dict[str, Tensor], momentum_buf: Dict[str, Tensor], x: Tensor, target: Tensor,
               variant: str = "egm", lr: float = 0.02, momentum: float = 0.9, wd: float = 1e-4,
               loss_weight=None):
    """fwd + criterion + bwd + SGD(momentum, wd) in place on `sd`. Returns (loss, grads)."""
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k]
    for k in names:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    upd = {}
    logits = forward(sd, x, variant, True, upd)
    loss = criterion(logits, target, loss_weight)
    loss.backward()
    grads = {}
    with torch.no_grad():
        for k in names:
            g = sd[k].grad
            grads[k] = g
            d = g + wd * sd[k]
            buf = momentum_buf.get(k)
            buf = d.clone() if buf is None else buf.mul_(momentum).add_(d)
            momentum_buf[k] = buf
            sd[k].sub_(lr * buf)
        for k, v in upd.items():
            sd[k] = v
    for k in names:
        sd[k].requires_grad_(False)
    return loss.detach(), grads
