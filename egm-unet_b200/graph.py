"""The EGM-UNet / UNet / yuanGRFBUNet graphs expressed over the tape engine (engine.py).

Each function mirrors one reference module's forward (file:line cited) and pushes the
hand-derived backward onto the tape.  Module objects are used purely as parameter
containers (same attribute tree as the reference, so state_dict keys match).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import abi
from .abi import call
from .engine import (ACT_NONE, ACT_RELU, ACT_SIGMOID, MODE_PLAIN, MODE_RESIDUAL, Ctx, SkipView, Var, WSpec, _p, bn_act, conv2d, conv_bn_act, conv_for_bn, conv_module,
                     copy_into, deconv_concat, edge_enhancer, from_nchw, maxpool2, mca_layer, out_conv, release_grad, slice_channels, to_nchw,
                     upsample_concat)


# =========================================================================== FusionConv (x1 is x2)
def fusion_conv(ctx: Ctx, cat: Var, m) -> Var:
    """src/EGM-UNet.py:1202-1236.  down(cat[x,x]) == conv1x1(x; W[:, :C] + W[:, C:]);  conv3+conv5+conv7 == one 7x7 conv
    with the three kernels summed (centre-embedded) and the three biases summed."""
    n, h, w, cc = cat.shape
    M, HW = n * h * w, h * w
    wdown = _p(m.down.weight)
    dim = wdown.shape[0]
    spec_down = WSpec(1, (m.down.weight,), (m.down.bias,), (dim, cc, 1, 1))
    spec_w7 = WSpec(2, (m.conv_7x7.weight, m.conv_5x5.weight, m.conv_3x3.weight), (m.conv_7x7.bias, m.conv_5x5.bias, m.conv_3x3.bias), (dim, dim, 7, 7))
    planned = ctx.wplan is not None and ctx.wplan.ready and ctx.record and spec_down.key in ctx.wplan.jobs and spec_w7.key in ctx.wplan.jobs

    def sink_b7(gb):
        for b in (m.conv_7x7.bias, m.conv_5x5.bias, m.conv_3x3.bias):
            call("copy_slice", gb, ctx.grad_slot(b), abi.F32, 1, dim, dim, 0, dim, 0, 0)

    if planned:      # folded / merged weights come packed out of egm_weight_prep_batch; gradients leave through egm_wgrad_unpack_batch
        f = conv2d(ctx, cat, None, m.down.bias, wspec=spec_down)
        s = conv2d(ctx, f, None, None, bgrad_sink=sink_b7, wspec=spec_w7)
    else:
        wfold = torch.empty(dim, cc, 1, 1, **ctx.f32)
        call("copy_slice", wdown, wfold, abi.F32, dim, cc, 2 * cc, 0, cc, 0, 0)
        call("copy_slice", wdown, wfold, abi.F32, dim, cc, 2 * cc, cc, cc, 0, 1)

        def sink_down(dw):
            g = ctx.grad_slot(m.down.weight)
            call("copy_slice", dw, g, abi.F32, dim, cc, cc, 0, 2 * cc, 0, 0)
            call("copy_slice", dw, g, abi.F32, dim, cc, cc, 0, 2 * cc, cc, 0)
        f = conv2d(ctx, cat, wfold, m.down.bias, wgrad_sink=sink_down, wspec=spec_down)

        w7 = torch.empty(dim, dim, 7, 7, **ctx.f32)
        call("kernel_embed", w7, _p(m.conv_7x7.weight), dim * dim, 7, 7, 0, 0)
        call("kernel_embed", w7, _p(m.conv_5x5.weight), dim * dim, 7, 5, 0, 1)
        call("kernel_embed", w7, _p(m.conv_3x3.weight), dim * dim, 7, 3, 0, 1)
        b7 = torch.empty(dim, **ctx.f32)
        call("copy_slice", _p(m.conv_7x7.bias), b7, abi.F32, 1, dim, dim, 0, dim, 0, 0)
        call("copy_slice", _p(m.conv_5x5.bias), b7, abi.F32, 1, dim, dim, 0, dim, 0, 1)
        call("copy_slice", _p(m.conv_3x3.bias), b7, abi.F32, 1, dim, dim, 0, dim, 0, 1)

        def sink_w7(dw):
            call("kernel_embed", dw, ctx.grad_slot(m.conv_7x7.weight), dim * dim, 7, 7, 1, 0)
            call("kernel_embed", dw, ctx.grad_slot(m.conv_5x5.weight), dim * dim, 7, 5, 1, 0)
            call("kernel_embed", dw, ctx.grad_slot(m.conv_3x3.weight), dim * dim, 7, 3, 1, 0)
        s = conv2d(ctx, f, w7, b7, wgrad_sink=sink_w7, bgrad_sink=sink_b7, wspec=spec_w7)

    # spatial attention on s (SpatialAttentionModule :1189-1200) and channel attention on f (ChannelAttentionModule :1171-1187) are
    # independent chains of small kernels, forward and backward: two branches of a parallel region (engine.Parallel)
    mm = torch.empty(M * 2, **ctx.f32)
    amax = torch.empty(M, dtype=torch.uint8, device=ctx.device)
    sa = torch.empty(M, **ctx.f32)
    wsa = _p(m.spatial_attention.conv1.weight)
    ca_m = m.channel_attention
    cr = ca_m.fc[0].weight.shape[0]
    avg, mx, ca = (torch.empty(n * dim, **ctx.f32) for _ in range(3))
    arg = torch.empty(n * dim, dtype=torch.int32, device=ctx.device)
    hid = torch.empty(2 * n * cr, **ctx.f32)
    w0, w2 = _p(ca_m.fc[0].weight), _p(ca_m.fc[2].weight)
    t = Var(ctx.empty(n, h, w, dim))
    hold = {}                          # dt: handed from the main-stream closure to the two branch closures in backward
    par = ctx.parallel(M)
    with par.branch(0):
        call("chan_meanmax", s.t, mm, amax, ctx.code, M, dim)
        call("sa_conv_fwd", mm, wsa, sa, n, h, w)
        if ctx.record:
            def bwd_sa():             # dsa -> dmm -> ds
                dt = hold.get("dt")
                if dt is None:
                    return
                dsa = torch.empty(M, **ctx.f32)
                call("pixel_dot", dt, s.t, ca, dsa, ctx.code, n, HW, dim)
                dmm = torch.empty(M * 2, **ctx.f32)
                gw = ctx.grad_slot(m.spatial_attention.conv1.weight)
                if ctx.wgrad_lane and planned:      # the 7x7 weight gradient is read by nobody before the optimizer: off the dependency chain
                    call("sa_conv_bwd", dsa, sa, mm, wsa, dmm, None, n, h, w)
                    ctx.wgrad_async(lambda: call("sa_conv_bwd", dsa, sa, mm, wsa, None, gw, n, h, w), (dsa, sa, mm))
                else:
                    call("sa_conv_bwd", dsa, sa, mm, wsa, dmm, gw, n, h, w)
                ds = ctx.empty(n, h, w, dim)
                call("fuse_mix_bwd_s", dt, sa, ca, dmm, amax, ds, ctx.code, n, HW, dim)
                s.accum(ds)
            ctx.push(bwd_sa)
    with par.branch(1):
        call("gap_gmp", f.t, avg, mx, arg, torch.empty(n * dim * 12 + 16, dtype=torch.uint8, device=ctx.device), ctx.code, n, HW, dim)
        call("ca_mlp_fwd", avg, mx, w0, w2, ca, hid, n, dim, cr)
        if ctx.record:
            def bwd_ca():             # channel-attention path + residual: df (+)= dt + davg/HW + [p == argmax] dmx
                dt = hold.get("dt")
                if dt is None:
                    return
                dca = torch.empty(n * dim, **ctx.f32)
                call("sample_chan_dot", dt, s.t, sa, dca, ctx.code, n, HW, dim)
                davg, dmx = torch.empty(n * dim, **ctx.f32), torch.empty(n * dim, **ctx.f32)
                call("ca_mlp_bwd", dca, ca, avg, mx, hid, w0, w2, ctx.grad_slot(ca_m.fc[0].weight), ctx.grad_slot(ca_m.fc[2].weight),
                     davg, dmx, n, dim, cr)
                gf, acc = f.grad_target()
                call("fuse_df_finish", gf, dt, davg, dmx, arg, acc, ctx.code, n, HW, dim)
            ctx.push(bwd_ca)
    par.join()
    call("fuse_mix_fwd", f.t, s.t, sa, ca, t.t, ctx.code, n, HW, dim)
    if ctx.record:
        def bwd_mix():                # runs first in backward (main stream): publish dt, then the two branch closures follow
            hold["dt"], t.grad = t.grad, None
        ctx.push(bwd_mix)
    return conv_module(ctx, t, m.up)


# =========================================================================== EdgeEnhancedGRFB
def basic_conv(ctx: Ctx, x: Var, m, out: Optional[Var] = None, out_coff: int = 0) -> Var:
    """BasicConv (src/EGM-UNet.py:958-975): conv -> BN(momentum 0.01) -> optional ReLU."""
    act = ACT_RELU if (m.relu is not None) else ACT_NONE
    return conv_bn_act(ctx, x, m.conv, m.bn, act, out=out, out_coff=out_coff)


def grfb(ctx: Ctx, x: Var, m) -> Var:
    """EdgeEnhancedGRFB.forward, src/EGM-UNet.py:1296-1323."""
    n, h, w, c = x.shape
    M = n * h * w
    ip = m.inter_planes
    xe = edge_enhancer(ctx, x, m.edge_enhancer)
    cat = Var(ctx.empty(n, h, w, c + 6 * ip))
    release_grad(ctx, cat)            # pushed first => runs after every slice consumer in backward
    copy_into(ctx, x, cat, 0)
    # The three branches are independent except that their first convs all back-propagate into xe's gradient.  Where the maps are small
    # (engine.Parallel) each branch runs on its own side stream, forward and backward; its first conv reads xe through a proxy Var
    # with a private gradient, and the three private gradients are summed into xe's after the join.  Every branch writes its own
    # channel slice of `cat` and only reads cat's gradient.
    par = ctx.parallel(M)
    first_in = par.on and ctx.grfb_first_in_region
    if first_in:
        xb = [Var(xe.t), Var(xe.t), Var(xe.t)]
        if ctx.record:
            def bwd_sum():            # pushed before the region => runs after it (main stream, behind the join)
                for v in xb:
                    g, v.grad = v.grad, None
                    if g is not None:
                        xe.accum(g)
            ctx.push(bwd_sum)
    else:
        xb = [xe, xe, xe]
    zs = None
    if first_in:                      # the shortcut's 1x1 conv only needs x: it joins branch 0; its BN (needs the fusion output) stays behind
        xs = Var(x.t)
        if ctx.record:
            def bwd_sum_x():
                g, xs.grad = xs.grad, None
                if g is not None:
                    x.accum(g)
            ctx.push(bwd_sum_x)
        with par.branch(0):
            zs, zs_sums = conv_for_bn(ctx, xs, m.shortcut.conv, m.shortcut.bn)
    if not first_in:                  # first convs on the main stream, one after the other (they accumulate into the same gradient)
        d0, e0, q0 = basic_conv(ctx, xe, m.branch_dir[0]), basic_conv(ctx, xe, m.branch_edge[0]), basic_conv(ctx, xe, m.branch_ctx[0])
    with par.branch(0):
        d = basic_conv(ctx, xb[0], m.branch_dir[0]) if first_in else d0
        d = basic_conv(ctx, d, m.branch_dir[1])
        basic_conv(ctx, d, m.branch_dir[2], out=cat, out_coff=c)
    with par.branch(1):
        e = basic_conv(ctx, xb[1], m.branch_edge[0]) if first_in else e0
        e = edge_enhancer(ctx, e, m.branch_edge[1])
        e = basic_conv(ctx, e, m.branch_edge[2])
        e = basic_conv(ctx, e, m.branch_edge[3])
        basic_conv(ctx, e, m.branch_edge[4], out=cat, out_coff=c + 2 * ip)
    with par.branch(2):
        q = basic_conv(ctx, xb[2], m.branch_ctx[0]) if first_in else q0
        q = basic_conv(ctx, q, m.branch_ctx[1])
        q = basic_conv(ctx, q, m.branch_ctx[2])
        basic_conv(ctx, q, m.branch_ctx[3], out=cat, out_coff=c + 4 * ip)
    par.join()
    fo = fusion_conv(ctx, cat, m.fusion_conv)
    if zs is not None:
        o = bn_act(ctx, zs, m.shortcut.bn, ACT_NONE, MODE_RESIDUAL, aux=fo, alpha=float(m.scale), sums=zs_sums)
    else:
        o = conv_bn_act(ctx, x, m.shortcut.conv, m.shortcut.bn, ACT_NONE, MODE_RESIDUAL, aux=fo, alpha=float(m.scale))
    tz = conv_module(ctx, o, m.target_enhancer[0])          # [N,H,W,3]
    y = Var(ctx.empty(n, h, w, o.C))
    call("mul_pixel_gate", o.t, tz.t, y.t, ctx.code, M, o.C, 3, 1)
    if ctx.record:
        def bwd():
            y._sync_grad()
            dy, y.grad = y.grad, None
            if dy is None:
                return
            dot = torch.empty(M, **ctx.f32)
            call("pixel_dot", dy, o.t, None, dot, ctx.code, n, h * w, o.C)
            dtz = ctx.empty(n, h, w, 3)
            call("pixel_gate_bwd", dot, tz.t, dtz, ctx.code, M, 3, 1)
            tz.accum(dtz)
            do = ctx.empty(n, h, w, o.C)
            call("mul_pixel_gate", dy, tz.t, do, ctx.code, M, o.C, 3, 1)
            o.accum(do)
        ctx.push(bwd)
    return y


# =========================================================================== RecursiveGatedAttention
def _unary(ctx: Ctx, x: Var, op_f: int, op_b: int, scalar: Optional[torch.Tensor] = None) -> Var:
    y = Var(torch.empty_like(x.t))
    nel = x.t.numel()
    call("unary", x.t, None, _p(scalar), y.t, ctx.code, nel, op_f)
    if ctx.record:
        def bwd():
            dy, y.grad = y.grad, None
            if dy is None:
                return
            dx = torch.empty_like(x.t)
            if op_f == 2:       # y = x * scalar
                call("dot_all", dy, x.t, ctx.grad_slot(scalar), ctx.code, nel)
                call("unary", dy, None, _p(scalar), dx, ctx.code, nel, 2)
            else:               # gelu
                call("unary", dy, x.t, None, dx, ctx.code, nel, op_b)
            x.accum(dx)
        ctx.push(bwd)
    return y


def _gate_mul(ctx: Ctx, a: Var, g: Var) -> Var:
    """a * sigmoid(g), g has one channel."""
    n, h, w, c = a.shape
    M = n * h * w
    y = Var(ctx.empty(n, h, w, c))
    call("mul_pixel_gate", a.t, g.t, y.t, ctx.code, M, c, 1, 0)
    if ctx.record:
        def bwd():
            dy, y.grad = y.grad, None
            if dy is None:
                return
            dot = torch.empty(M, **ctx.f32)
            call("pixel_dot", dy, a.t, None, dot, ctx.code, n, h * w, c)
            dg = ctx.empty(n, h, w, 1)
            call("pixel_gate_bwd", dot, g.t, dg, ctx.code, M, 1, 0)
            g.accum(dg)
            da = ctx.empty(n, h, w, c)
            call("mul_pixel_gate", dy, g.t, da, ctx.code, M, c, 1, 0)
            a.accum(da)
        ctx.push(bwd)
    return y


def rga(ctx: Ctx, x: Var, m) -> Var:
    """RecursiveGatedAttention.forward, src/EGM-UNet.py:518-547."""
    assert m.order == 2
    s0, stot = m.split_sizes[0], sum(m.split_sizes)
    fused = conv_module(ctx, x, m.proj_in)
    base = slice_channels(ctx, fused, 0, s0)
    gconv = conv_module(ctx, fused, m.dwconv, x_coff=s0, x_cin=stot)
    gates = _unary(ctx, gconv, 2, 2, scalar=m.scale)
    out = base
    off = 0
    for i in range(m.order):
        ci = m.split_sizes[i]
        g = conv_module(ctx, gates, m.gate_convs[i][0], x_coff=off, x_cin=ci)
        g = _unary(ctx, g, 0, 1)
        g = conv_module(ctx, g, m.gate_convs[i][2])
        out = _gate_mul(ctx, out, g)
        if i < m.order - 1:
            out = conv_module(ctx, out, m.transform_convs[i])
        off += ci
    return conv_module(ctx, out, m.proj_out)


# =========================================================================== skeleton
def double_conv(ctx: Ctx, x: Var, seq, i0: int = 0, i1: int = 3, skip_into: Optional[Var] = None):
    """DoubleConv, src/EGM-UNet.py:44-55 == src/unet.py:7-18.  skip_into: the Up level's concat buffer -- the result is written into its
    first channels and returned as a SkipView (virtual concat)."""
    x = conv_bn_act(ctx, x, seq[i0], seq[i0 + 1], ACT_RELU)
    if skip_into is None:
        return conv_bn_act(ctx, x, seq[i1], seq[i1 + 1], ACT_RELU)
    conv_bn_act(ctx, x, seq[i1], seq[i1 + 1], ACT_RELU, out=skip_into, out_coff=0)
    return SkipView(skip_into, seq[i1].out_channels)


def down_block(ctx: Ctx, x, down, variant: str, skip_into: Optional[Var] = None):
    """Down: src/unet.py:21-26 ('unet'), src/EGM-UNet.py:888-912 ('egm'), src/yuanGRFBUNet.py:859-883 ('yuan')."""
    x = maxpool2(ctx, x)
    seq = down[1]
    if variant == "unet":
        return double_conv(ctx, x, seq, skip_into=skip_into)
    x = conv_bn_act(ctx, x, seq[0], seq[1], ACT_RELU)
    if variant == "egm":
        x = mca_layer(ctx, x, seq[3])
        c2, gi = 4, 7
    else:
        c2, gi = 3, 6
    x = conv_bn_act(ctx, x, seq[c2], seq[c2 + 1], ACT_RELU)
    return grfb(ctx, x, seq[gi])


def up_block(ctx: Ctx, low: Var, skip, up) -> Var:
    """Up (bilinear): src/EGM-UNet.py:927-949 == src/unet.py:29-51."""
    if isinstance(up.up, nn.Upsample):
        return double_conv(ctx, upsample_concat(ctx, low, skip), up.conv)
    return double_conv(ctx, deconv_concat(ctx, low, skip, up.up), up.conv)


def _skip_buffer(ctx: Ctx, n: int, h: int, w: int, cs: int, up, producer_ok: bool) -> Optional[Var]:
    """Concat buffer of one Up level, allocated when its skip connection is PRODUCED (virtual concat), or None where the skip's producer
    cannot write into a channel slice (the GRFB tail of the EGM / yuan Down blocks) or Up is a transposed conv."""
    if not (ctx.virtual_skip and producer_ok and isinstance(up.up, nn.Upsample)):
        return None
    cat = Var(ctx.empty(n, h, w, up.conv[0].in_channels))
    assert cat.C > cs
    release_grad(ctx, cat)            # pushed before every user => runs after all of them in backward
    return cat


def net_forward(ctx: Ctx, model, x_nchw: torch.Tensor, variant: str):
    """GRFBUNet.forward (src/EGM-UNet.py:1527-1541) / UNet.forward (src/unet.py:84-96).
    Returns (logits NCHW fp32, logits Var)."""
    x = from_nchw(ctx, x_nchw)
    n, h, w, _ = x.shape
    plain = variant == "unet"
    x1 = double_conv(ctx, x, model.in_conv, skip_into=_skip_buffer(ctx, n, h, w, model.in_conv[3].out_channels, model.up4, True))
    x2 = down_block(ctx, x1, model.down1, variant, _skip_buffer(ctx, n, h // 2, w // 2, model.down1[1][3].out_channels, model.up3, plain) if plain else None)
    x3 = down_block(ctx, x2, model.down2, variant, _skip_buffer(ctx, n, h // 4, w // 4, model.down2[1][3].out_channels, model.up2, plain) if plain else None)
    x4 = down_block(ctx, x3, model.down3, variant, _skip_buffer(ctx, n, h // 8, w // 8, model.down3[1][3].out_channels, model.up1, plain) if plain else None)
    x5 = down_block(ctx, x4, model.down4, variant)
    if variant != "unet":
        x5 = rga(ctx, x5, model.attn1)
    y = up_block(ctx, x5, x4, model.up1)
    y = up_block(ctx, y, x3, model.up2)
    y = up_block(ctx, y, x2, model.up3)
    y = up_block(ctx, y, x1, model.up4)
    return out_conv(ctx, y, model.out_conv[0])          # (fp32 NCHW logits, backward-seed handle)
