"""Drop-in `UNet` / `GRFBUNet` (EGM-UNet) / yuan `GRFBUNet` modules.

Constructors, attribute trees, parameter names/shapes and `forward(x) -> {"out": logits}`
are those of the reference (src/unet.py:61-96, src/EGM-UNet.py:1503-1541,
src/yuanGRFBUNet.py:1474-1512), so `state_dict()` / `load_state_dict()` interchange with
reference checkpoints and train.py / predict.py run unchanged.  The torch layer objects
inside are only parameter containers: `forward` hands the whole graph to the tape engine,
which runs it as libegm_b200 CUDA kernels (no cuDNN / cuBLAS / ATen compute, no CPU path).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import abi
from .abi import call
from .engine import Ctx, seed_grad_from_nchw
from .graph import net_forward


# --------------------------------------------------------------------------- containers (reference-identical trees)
class DoubleConv(nn.Sequential):
    """src/unet.py:7-18 / src/EGM-UNet.py:44-55."""

    def __init__(self, in_channels, out_channels, mid_channels=None, use_attention=False):
        if mid_channels is None:
            mid_channels = out_channels
        super().__init__(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))


class MCAGate(nn.Module):
    """src/EGM-UNet.py:836-869 (pools hold no parameters)."""

    def __init__(self, k_size):
        super().__init__()
        self.conv = nn.Conv2d(1, 1, kernel_size=(1, k_size), stride=1, padding=(0, (k_size - 1) // 2), bias=False)
        self.weight = nn.Parameter(torch.rand(2))


class MCALayer(nn.Module):
    """src/EGM-UNet.py:686-705."""

    def __init__(self, inp):
        super().__init__()
        self.inp = inp
        temp = round(abs((math.log2(inp) - 1) / 1.5))
        kernel = temp if temp % 2 else temp - 1
        self.h_cw = MCAGate(3)
        self.w_hc = MCAGate(3)
        self.c_hw = MCAGate(kernel)


class EdgeAwareFeatureEnhancer(nn.Module):
    """src/EGM-UNet.py:872-886."""

    def __init__(self, in_channels):
        super().__init__()
        self.weight_generator = nn.Sequential(nn.Conv2d(in_channels, in_channels, kernel_size=1), nn.BatchNorm2d(in_channels), nn.Sigmoid())


class BasicConv(nn.Module):
    """src/EGM-UNet.py:958-975."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu=True, bn=True, bias=False):
        super().__init__()
        self.out_channels = out_channels
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation,
                              groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None


class ChannelAttentionModule(nn.Module):
    """src/EGM-UNet.py:1171-1187."""

    def __init__(self, in_channels, reduction=4):
        super().__init__()
        self.fc = nn.Sequential(nn.Conv2d(in_channels, in_channels // reduction, 1, bias=False), nn.ReLU(inplace=True),
                                nn.Conv2d(in_channels // reduction, in_channels, 1, bias=False))


class SpatialAttentionModule(nn.Module):
    """src/EGM-UNet.py:1189-1200."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)


class FusionConv(nn.Module):
    """src/EGM-UNet.py:1202-1219."""

    def __init__(self, in_channels, out_channels, factor=4.0):
        super().__init__()
        dim = int(out_channels // factor)
        self.down = nn.Conv2d(2 * in_channels, dim, kernel_size=1, stride=1)
        self.conv_3x3 = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1)
        self.conv_5x5 = nn.Conv2d(dim, dim, kernel_size=5, stride=1, padding=2)
        self.conv_7x7 = nn.Conv2d(dim, dim, kernel_size=7, stride=1, padding=3)
        self.spatial_attention = SpatialAttentionModule()
        self.channel_attention = ChannelAttentionModule(dim)
        self.up = nn.Conv2d(dim, out_channels, kernel_size=1, stride=1)


class EdgeEnhancedGRFB(nn.Module):
    """src/EGM-UNet.py:1238-1294."""

    def __init__(self, in_channels, out_channels, stride=1, scale=0.1, visual=12, fusion_factor=4.0):
        super().__init__()
        self.scale = scale
        self.out_channels = out_channels
        ip = self.inter_planes = max(in_channels // 8, 4)
        self.edge_enhancer = EdgeAwareFeatureEnhancer(in_channels)
        self.branch_dir = nn.Sequential(
            BasicConv(in_channels, 2 * ip, 1),
            BasicConv(2 * ip, 2 * ip, 3, padding=visual, dilation=visual, relu=False),
            BasicConv(2 * ip, 2 * ip, 1))
        self.branch_edge = nn.Sequential(
            BasicConv(in_channels, ip, 1),
            EdgeAwareFeatureEnhancer(ip),
            BasicConv(ip, 2 * ip, (3, 3), stride, padding=1, groups=ip),
            BasicConv(2 * ip, 2 * ip, 3, padding=2 * visual, dilation=2 * visual, relu=False),
            BasicConv(2 * ip, 2 * ip, 1))
        self.branch_ctx = nn.Sequential(
            BasicConv(in_channels, ip, 3, padding=1),
            BasicConv(ip, 2 * ip, 3, stride=stride, padding=1, groups=2),
            BasicConv(2 * ip, 2 * ip, 3, padding=3 * visual, dilation=3 * visual, relu=False),
            BasicConv(2 * ip, 2 * ip, 1))
        self.concat_channels = in_channels + 6 * ip
        self.fusion_conv = FusionConv(self.concat_channels, out_channels, factor=fusion_factor)
        self.shortcut = BasicConv(in_channels, out_channels, 1, stride, relu=False)
        self.relu = nn.ReLU(inplace=False)
        self.target_enhancer = nn.Sequential(nn.Conv2d(out_channels, 3, 3, padding=1), nn.Sigmoid())


class DoubleConv1(nn.Sequential):
    """src/EGM-UNet.py:888-904 (with MCALayer) / src/yuanGRFBUNet.py:859-875 (without)."""

    def __init__(self, in_channels, out_channels, mid_channels=None, with_mca=True):
        if mid_channels is None:
            mid_channels = out_channels
        layers = [nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True)]
        if with_mca:
            layers.append(MCALayer(mid_channels))
        layers += [nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
                   EdgeEnhancedGRFB(mid_channels, out_channels, stride=1, scale=0.1, visual=12)]
        super().__init__(*layers)


class Down(nn.Sequential):
    def __init__(self, in_channels, out_channels, kind="unet"):
        body = DoubleConv(in_channels, out_channels) if kind == "unet" else DoubleConv1(in_channels, out_channels, with_mca=(kind == "egm"))
        super().__init__(nn.MaxPool2d(2, stride=2), body)


class Up(nn.Module):
    """src/unet.py:29-37 / src/EGM-UNet.py:927-936."""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)


class OutConv(nn.Sequential):
    def __init__(self, in_channels, num_classes):
        super().__init__(nn.Conv2d(in_channels, num_classes, kernel_size=1))


class RecursiveGatedAttention(nn.Module):
    """src/EGM-UNet.py:458-516."""

    def __init__(self, dim, order=2, reduction=8, kernel_size=3):
        super().__init__()
        self.order, self.dim = order, dim
        self.split_sizes = [dim // (2 ** i) for i in range(1, order)]
        self.split_sizes.append(dim // (2 ** (order - 1)))
        self.split_sizes.reverse()
        total = sum(self.split_sizes)
        if total > dim:
            self.split_sizes[-1] = dim - sum(self.split_sizes[:-1])
        self.proj_in = nn.Conv2d(dim, self.split_sizes[0] + sum(self.split_sizes), 1)
        self.gate_convs = nn.ModuleList()
        for i in range(order):
            in_ch = self.split_sizes[i]
            self.gate_convs.append(nn.Sequential(nn.Conv2d(in_ch, max(in_ch // reduction, 8), 1), nn.GELU(),
                                                 nn.Conv2d(max(in_ch // reduction, 8), 1, 1), nn.Sigmoid()))
        self.transform_convs = nn.ModuleList()
        for i in range(order - 1):
            self.transform_convs.append(nn.Conv2d(self.split_sizes[i], self.split_sizes[i + 1], 1))
        self.dwconv = nn.Conv2d(sum(self.split_sizes), sum(self.split_sizes), kernel_size, padding=kernel_size // 2, groups=sum(self.split_sizes))
        self.proj_out = nn.Conv2d(self.split_sizes[-1], dim, 1)
        self.scale = nn.Parameter(torch.tensor(1.0))
        print(f"[RGA] order={order}, split_sizes={self.split_sizes}")


# --------------------------------------------------------------------------- parameter storage
class ParamStore:
    """Flat fp32 parameter / gradient / momentum buffers; every nn.Parameter becomes a view into `params`, and the kernels
    write dL/dp into the matching view of `grads` (one contiguous bucket for the all-reduce and the fused SGD)."""

    ALIGN = 8       # floats: keeps every slot 32-byte aligned for vector loads

    def __init__(self, model: nn.Module):
        self.model = model
        self.plist = [p for p in model.parameters()]
        self.offsets, off = [], 0
        for p in self.plist:
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        dev = self.plist[0].device
        self.params = torch.empty(off, dtype=torch.float32, device=dev)
        self.grads = torch.empty(off, dtype=torch.float32, device=dev)
        call("memset_zero", self.params, off * 4)
        call("memset_zero", self.grads, off * 4)
        self.momentum: Optional[torch.Tensor] = None
        self._gview: Dict[int, torch.Tensor] = {}
        for p, o in zip(self.plist, self.offsets):
            n = p.numel()
            v = self.params[o:o + n].view(p.shape)
            call("copy_slice", p.data.contiguous(), v, abi.F32, 1, n, n, 0, n, 0, 0)
            p.data = v
            self._gview[id(p)] = self.grads[o:o + n].view(p.shape)
        self._ptrs = [p.data_ptr() for p in self.plist]
        self.index = {id(p): i for i, p in enumerate(self.plist)}
        self.on_grad = None          # optional callback(param_index) -- the DDP bucket reducer hooks in here
        self.ranges = [(o, (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN) for p, o in zip(self.plist, self.offsets)]

    def valid(self) -> bool:
        return all(p.data_ptr() == q for p, q in zip(self.plist, self._ptrs))

    def grad_slot(self, p: torch.Tensor) -> torch.Tensor:
        if self.on_grad is not None:
            self.on_grad(self.index[id(p)])
        return self._gview[id(p)]


def _store_for(model: nn.Module) -> ParamStore:
    st = getattr(model, "_egm_store", None)
    if st is None or not st.valid():
        st = ParamStore(model)
        object.__setattr__(model, "_egm_store", st)
    return st


# --------------------------------------------------------------------------- autograd bridge
class _NetFn(torch.autograd.Function):
    """Whole-network forward/backward as ONE autograd node: forward runs the tape engine, backward replays the tape and
    returns the parameter gradients (fresh copies of the flat gradient bucket views)."""

    @staticmethod
    def forward(fctx, model, x, *params):
        store = _store_for(model)
        need = any(fctx.needs_input_grad)      # (grad mode is off inside Function.forward)
        ctx = Ctx(model.compute_dtype, x.device, model.training, need, store.grad_slot, use_tc=model.use_tensor_cores)
        logits, lv = net_forward(ctx, model, x.detach().float(), model.variant)
        fctx.ectx, fctx.lv, fctx.store, fctx.nparams = ctx, lv, store, len(params)
        return logits

    @staticmethod
    def backward(fctx, dlogits):
        ctx, store = fctx.ectx, fctx.store
        seed_grad_from_nchw(ctx, fctx.lv, dlogits.float())
        ctx.backward()
        outs = []
        for p in store.plist:
            g = torch.empty_like(p.data)
            n = g.numel()
            call("copy_slice", store.grad_slot(p), g, abi.F32, 1, n, n, 0, n, 0, 0)
            outs.append(g)
        fctx.ectx = fctx.lv = None
        return (None, None, *outs)


class _B200Net(nn.Module):
    variant = "unet"

    def _init_runtime(self):
        object.__setattr__(self, "_egm_store", None)
        self.compute_dtype = torch.bfloat16      # production dtype; set_check_mode(True) -> fp32 end to end
        self.use_tensor_cores = True

    def set_check_mode(self, fp32: bool = True):
        """fp32 check mode (BASELINE.json north_star: logits within 1e-5 of the reference)."""
        self.compute_dtype = torch.float32 if fp32 else torch.bfloat16
        return self

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        if not x.is_cuda:
            raise RuntimeError("egm_b200 models run on a B200 (sm_100a) only; there is no CPU path (the CPU oracle lives in oracle/)")
        params = [p for p in self.parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            logits = _NetFn.apply(self, x, *params)
        else:
            store = _store_for(self)
            ctx = Ctx(self.compute_dtype, x.device, self.training, False, store.grad_slot, use_tc=self.use_tensor_cores)
            logits, _ = net_forward(ctx, self, x.detach().float(), self.variant)
        return {"out": logits}


class UNet(_B200Net):
    """src/unet.py:61-96."""
    variant = "unet"

    def __init__(self, in_channels: int = 1, num_classes: int = 2, bilinear: bool = True, base_c: int = 64):
        super().__init__()
        self.in_channels, self.num_classes, self.bilinear = in_channels, num_classes, bilinear
        self.in_conv = DoubleConv(in_channels, base_c)
        self.down1 = Down(base_c, base_c * 2)
        self.down2 = Down(base_c * 2, base_c * 4)
        self.down3 = Down(base_c * 4, base_c * 8)
        factor = 2 if bilinear else 1
        self.down4 = Down(base_c * 8, base_c * 16 // factor)
        self.up1 = Up(base_c * 16, base_c * 8 // factor, bilinear)
        self.up2 = Up(base_c * 8, base_c * 4 // factor, bilinear)
        self.up3 = Up(base_c * 4, base_c * 2 // factor, bilinear)
        self.up4 = Up(base_c * 2, base_c, bilinear)
        self.out_conv = OutConv(base_c, num_classes)
        self._init_runtime()


class GRFBUNet(_B200Net):
    """EGM-UNet: src/EGM-UNet.py:1503-1541."""
    variant = "egm"

    def __init__(self, in_channels: int = 1, num_classes: int = 2, bilinear: bool = True, base_c: int = 64, use_attention: bool = False):
        super().__init__()
        if not bilinear:
            # the reference itself cannot build this (src/EGM-UNet.py:935 passes use_attention as mid_channels -> TypeError)
            raise TypeError("GRFBUNet(bilinear=False) is not constructible in the reference either (src/EGM-UNet.py:935)")
        self.in_channels, self.num_classes, self.bilinear = in_channels, num_classes, bilinear
        kind = self.variant
        self.in_conv = DoubleConv(in_channels, base_c)
        self.down1 = Down(base_c, base_c * 2, kind)
        self.down2 = Down(base_c * 2, base_c * 4, kind)
        self.down3 = Down(base_c * 4, base_c * 8, kind)
        factor = 2
        self.down4 = Down(base_c * 8, base_c * 16 // factor, kind)
        self.attn1 = RecursiveGatedAttention(base_c * 16 // factor)
        self.up1 = Up(base_c * 16, base_c * 8 // factor, bilinear)
        self.up2 = Up(base_c * 8, base_c * 4 // factor, bilinear)
        self.up3 = Up(base_c * 4, base_c * 2 // factor, bilinear)
        self.up4 = Up(base_c * 2, base_c, bilinear)
        self.out_conv = OutConv(base_c, num_classes)
        self._init_runtime()


class YuanGRFBUNet(GRFBUNet):
    """src/yuanGRFBUNet.py:1474-1512 (DoubleConv1 without MCALayer)."""
    variant = "yuan"
