"""GPU: every engine op (forward + hand-written backward) against torch CPU autograd on the same inputs, in fp32 check mode
(tolerance 2e-4 rel-to-max) and bf16 production mode (4e-2)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import egm_oracle as O
from tests.util import Harness, TOL, rel_err, cosine

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]


def _rand(*s, seed=0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed))


def _q(x, dtype):
    """round inputs to the storage dtype so both sides see identical values"""
    return x.to(dtype).float()


CONV_CASES = [  # cin, cout, k, dil, groups, bias, H, W
    (3, 32, 3, 1, 1, False, 20, 24), (32, 64, 3, 1, 1, False, 17, 19), (64, 16, 1, 1, 1, False, 12, 12), (16, 16, 3, 12, 1, False, 30, 26),
    (8, 16, 3, 1, 8, False, 14, 14), (8, 16, 3, 1, 2, False, 14, 14), (16, 16, 7, 1, 1, True, 15, 13), (64, 3, 3, 1, 1, True, 10, 10),
    (16, 1, 1, 1, 1, True, 9, 9), (32, 2, 1, 1, 1, True, 11, 7), (24, 24, 3, 1, 24, True, 9, 10), (112, 16, 1, 1, 1, True, 8, 8),
    (16, 16, 3, 36, 1, False, 30, 30)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_direct(case, dtype):
    from egm_unet_b200.engine import conv_module
    cin, cout, k, dil, groups, bias, H, W = case
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, groups=groups, bias=bias)
    x = _q(_rand(2, cin, H, W), dtype)
    g = _q(_rand(2, cout, H, W, seed=3), dtype)
    hs = Harness(dtype, use_tc=False)
    xv = hs.var(x)
    mc = m.cuda()
    yv = conv_module(hs.ctx, xv, mc)
    y = hs.out(yv)
    xr = x.clone().requires_grad_(True)
    m = m.cpu()
    yr = m(xr)
    yr.backward(g)
    tol = TOL[dtype]
    assert rel_err(y, yr.detach()) < tol
    hs.backward(yv, g)
    assert rel_err(hs.grad(xv), xr.grad) < tol
    assert rel_err(hs.pgrad(mc.weight), m.weight.grad) < tol
    if bias:
        assert rel_err(hs.pgrad(mc.bias), m.bias.grad) < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["relu", "none", "edge", "residual", "eval"])
def test_bn_act(mode, dtype):
    from egm_unet_b200 import engine as E
    C, H, W = 24, 13, 11
    torch.manual_seed(2)
    bn = nn.BatchNorm2d(C, momentum=0.01 if mode == "none" else 0.1)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C)); bn.bias.copy_(0.1 * torch.randn(C))
        bn.running_mean.copy_(0.1 * torch.randn(C)); bn.running_var.copy_(1 + 0.2 * torch.rand(C))
    import copy
    bn_ref = copy.deepcopy(bn)
    z = _q(_rand(3, C, H, W) * 1.5 + 0.3, dtype)
    aux = _q(_rand(3, C, H, W, seed=5), dtype)
    g = _q(_rand(3, C, H, W, seed=7), dtype)
    training = mode != "eval"
    hs = Harness(dtype, training=training)
    zv, av = hs.var(z), hs.var(aux)
    bnc = bn.cuda()
    if mode in ("relu", "eval"):
        yv = E.bn_act(hs.ctx, zv, bnc, E.ACT_RELU)
    elif mode == "none":
        yv = E.bn_act(hs.ctx, zv, bnc, E.ACT_NONE)
    elif mode == "edge":
        yv = E.bn_act(hs.ctx, zv, bnc, E.ACT_SIGMOID, E.MODE_EDGE_GATE, aux=av)
    else:
        yv = E.bn_act(hs.ctx, zv, bnc, E.ACT_NONE, E.MODE_RESIDUAL, aux=av, alpha=0.1)
    y = hs.out(yv)
    zr, ar = z.clone().requires_grad_(True), aux.clone().requires_grad_(True)
    bn_ref.train(training)
    b = bn_ref(zr)
    yr = {"relu": lambda: F.relu(b), "eval": lambda: F.relu(b), "none": lambda: b, "edge": lambda: torch.sigmoid(b) * ar + ar,
          "residual": lambda: F.relu(0.1 * ar + b)}[mode]()
    yr.backward(g)
    tol = TOL[dtype]
    assert rel_err(y, yr.detach()) < tol
    hs.backward(yv, g)
    assert rel_err(hs.grad(zv), zr.grad) < tol * 3
    assert rel_err(hs.pgrad(bnc.weight), bn_ref.weight.grad) < tol * 3
    assert rel_err(hs.pgrad(bnc.bias), bn_ref.bias.grad) < tol * 3
    if mode in ("edge", "residual"):
        assert rel_err(hs.grad(av), ar.grad) < tol
    if training:
        assert rel_err(bnc.running_mean.cpu(), bn_ref.running_mean) < 1e-4 if dtype == torch.float32 else True
        assert rel_err(bnc.running_var.cpu(), bn_ref.running_var) < 1e-4 if dtype == torch.float32 else True
        assert int(bnc.num_batches_tracked) == 1


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw", [(12, 16), (13, 15)])
def test_maxpool(hw, dtype):
    from egm_unet_b200 import engine as E
    x = _q(_rand(2, 16, *hw), dtype)
    x[:, :, 2:6, 2:6] = 0.0          # ties: first-max rule
    hs = Harness(dtype)
    xv = hs.var(x)
    yv = E.maxpool2(hs.ctx, xv)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    g = _q(_rand(*yr.shape, seed=9), dtype)
    yr.backward(g)
    assert rel_err(hs.out(yv), yr.detach()) < 1e-6
    hs.backward(yv, g)
    assert rel_err(hs.grad(xv), xr.grad) < 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [((6, 7), (12, 14)), ((6, 7), (13, 15)), ((5, 5), (11, 10))])
def test_upsample_concat(shape, dtype):
    from egm_unet_b200 import engine as E
    (hl, wl), (h, w) = shape
    low, skip = _q(_rand(2, 16, hl, wl), dtype), _q(_rand(2, 8, h, w, seed=4), dtype)
    hs = Harness(dtype)
    lv, sv = hs.var(low), hs.var(skip)
    ov = E.upsample_concat(hs.ctx, lv, sv)
    lr, sr = low.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    up = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = h - up.shape[2], w - up.shape[3]
    ref = torch.cat([sr, F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])], 1)
    g = _q(_rand(*ref.shape, seed=6), dtype)
    ref.backward(g)
    tol = TOL[dtype]
    assert rel_err(hs.out(ov), ref.detach()) < tol
    hs.backward(ov, g)
    assert rel_err(hs.grad(lv), lr.grad) < tol
    assert rel_err(hs.grad(sv), sr.grad) < tol


def _wrap(mod):
    return nn.ModuleDict({"m": mod})


def _run_block(fn_engine, fn_oracle, mod, x, dtype, training=True, tol_scale=1.0, grad_cos=0.999):
    """forward + backward of one block on both sides; returns nothing, asserts parity"""
    import copy
    from oracle import synth
    wrapped = _wrap(mod)
    sd = synth.fill_state_dict(wrapped.state_dict())
    wrapped.load_state_dict(sd)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    names = [k for k, _ in wrapped.named_parameters()]
    for k in names:
        ref_sd[k].requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    yr = fn_oracle(ref_sd, "m", xr)
    g = _q(_rand(*yr.shape, seed=11), dtype)
    yr.backward(g)
    hs = Harness(dtype, training=training, use_tc=False)
    wrapped.cuda()
    xv = hs.var(x)
    yv = fn_engine(hs.ctx, xv, wrapped["m"])
    tol = TOL[dtype] * tol_scale
    e = rel_err(hs.out(yv), yr.detach())
    assert e < tol, f"forward rel err {e}"
    hs.backward(yv, g)
    e = rel_err(hs.grad(xv), xr.grad)
    assert e < tol * 5, f"input grad rel err {e}"
    bad = []
    if dtype == torch.bfloat16:
        grad_cos = min(grad_cos, 0.98)
    gmax = max(float(ref_sd[k].grad.norm()) for k in names if ref_sd[k].grad is not None)
    for k, p in wrapped.named_parameters():
        gr = ref_sd[k].grad
        # parameters shifted away by a following train-mode BN have an analytically ZERO gradient (pure round-off): skip
        if gr is None or float(gr.norm()) < 1e-5 * gmax:
            continue
        got = hs.pgrad(p)
        if got.numel() > 4:
            ok = cosine(got, gr) > grad_cos and abs(float(got.norm()) / float(gr.norm()) - 1) < (0.05 if dtype == torch.float32 else 0.15)
        else:
            ok = rel_err(got, gr) < max(tol * 20, 2e-3)
        if not ok:
            bad.append((k, rel_err(got, gr), cosine(got, gr)))
    assert not bad, bad


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,hw", [(64, (12, 10)), (256, (7, 9)), (16, (9, 9)), (64, (70, 61)), (128, (33, 40)), (64, (5, 29))])
def test_mca_layer(C, hw, dtype):
    from egm_unet_b200 import engine as E
    from egm_unet_b200.models import MCALayer
    x = _q(F.relu(_rand(2, C, *hw)), dtype)           # post-ReLU input, many exact zeros -> exercises the tie rules
    _run_block(lambda c, v, m: E.mca_layer(c, v, m), lambda sd, p, t: O.mca_layer(sd, p, t), MCALayer(C), x, dtype, tol_scale=2.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,hw", [(16, (11, 13)), (32, (35, 29)), (64, (17, 40)), (128, (9, 10))])
def test_edge_enhancer(C, hw, dtype):
    """src/EGM-UNet.py:872-886.  bf16, C <= 64: high-pass + 1x1 conv run as ONE composed 3x3 tcgen05 conv (statistics in its epilogue);
    fp32 and C = 128: row-walking high-pass kernel + 1x1 conv.  Both against the oracle (forward, dx, every parameter gradient)."""
    from egm_unet_b200 import engine as E
    from egm_unet_b200.models import EdgeAwareFeatureEnhancer
    x = _q(_rand(2, C, *hw), dtype)
    _run_block(lambda c, v, m: E.edge_enhancer(c, v, m), lambda sd, p, t: O.edge_enhancer(sd, p, t, True, None), EdgeAwareFeatureEnhancer(C), x, dtype)


@pytest.mark.parametrize("C,hw", [(16, (24, 19)), (64, (33, 47))])
def test_edge_enhancer_composed_conv_equals_highpass_then_conv(C, hw):
    """The composed 3x3 conv (egm_highpass_compose) against the unfused bf16 path on the same inputs: same function, the unfused path
    additionally rounds the high-pass tensor to bf16, so the two agree to bf16 rounding; a constant image gives exactly zero pre-BN
    response in the composed form (centre weight == -8 x off-centre weight in bf16)."""
    from egm_unet_b200 import engine as E
    from egm_unet_b200.models import EdgeAwareFeatureEnhancer
    from egm_unet_b200.abi import call
    torch.manual_seed(5)
    m = EdgeAwareFeatureEnhancer(C).cuda()
    x = _q(_rand(2, C, *hw), torch.bfloat16)
    g = _q(_rand(2, C, *hw, seed=9), torch.bfloat16)
    res = []
    for fuse in (True, False):
        hs = Harness(torch.bfloat16, use_tc=True)
        hs.ctx.fuse_edge = fuse
        xv = hs.var(x)
        yv = E.edge_enhancer(hs.ctx, xv, m)
        y = hs.out(yv)
        hs.backward(yv, g)
        res.append((y, hs.grad(xv), hs.pgrad(m.weight_generator[0].weight).clone(), hs.pgrad(m.weight_generator[1].weight).clone()))
    for a, b in zip(*res):
        assert rel_err(a, b) < 2e-2, rel_err(a, b)
    # exact zero response of the composed weight to a constant input
    w1 = torch.randn(C, C, device="cuda")
    w3 = torch.empty(C, C, 3, 3, device="cuda")
    call("highpass_compose", w1, w3, C * C, 0, 1)
    torch.cuda.synchronize()
    assert torch.equal(w3.bfloat16().float(), w3) and float(w3.double().sum(dim=(2, 3)).abs().max()) == 0.0
    dw1 = torch.empty(C, C, device="cuda")
    call("highpass_compose", dw1, w3, C * C, 1, 0)
    torch.cuda.synchronize()
    k = torch.full((3, 3), -1 / 9.0, device="cuda"); k[1, 1] = 8 / 9.0
    assert rel_err(dw1.cpu(), (w3 * k).sum(dim=(2, 3)).cpu()) < 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
def test_fusion_conv(dtype):
    from egm_unet_b200 import graph as G
    from egm_unet_b200.models import FusionConv
    x = _q(_rand(2, 40, 10, 12), dtype)
    _run_block(lambda c, v, m: G.fusion_conv(c, v, m), lambda sd, p, t: O.fusion_conv(sd, p, t), FusionConv(40, 32), x, dtype, tol_scale=2.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,hw", [(64, (16, 14)), (128, (9, 8))])
def test_grfb(C, hw, dtype):
    from egm_unet_b200 import graph as G
    from egm_unet_b200.models import EdgeEnhancedGRFB
    x = _q(F.relu(_rand(2, C, *hw)), dtype)
    _run_block(lambda c, v, m: G.grfb(c, v, m), lambda sd, p, t: O.grfb(sd, p, t, True, None), EdgeEnhancedGRFB(C, C), x, dtype,
               tol_scale=3.0, grad_cos=0.995)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rga(dtype):
    from egm_unet_b200 import graph as G
    from egm_unet_b200.models import RecursiveGatedAttention
    x = _q(_rand(2, 64, 6, 7), dtype)
    _run_block(lambda c, v, m: G.rga(c, v, m), lambda sd, p, t: O.rga(sd, p, t), RecursiveGatedAttention(64), x, dtype, tol_scale=2.0)


@pytest.mark.parametrize("shape", [(2, 2, 24, 20), (3, 2, 17, 33), (2, 3, 16, 16)])
def test_fused_loss(shape):
    import egm_unet_b200 as EG
    from oracle import synth
    n, c, h, w = shape
    logits = _rand(n, c, h, w) * 2
    target = torch.randint(0, c, (n, h, w), generator=torch.Generator().manual_seed(5))
    target[:, :3] = 255
    lw = torch.tensor([1.0, 2.0, 0.5][:c])
    lr = logits.clone().requires_grad_(True)
    ref = O.criterion(lr, target, lw, num_classes=c)
    ref.backward()
    lg = logits.cuda().requires_grad_(True)
    loss = EG.criterion({"out": lg}, target.cuda(), lw.cuda(), num_classes=c, ignore_index=255)
    (loss * 1.0).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(lg.grad.cpu(), lr.grad) < 1e-4
    terms = O.loss_terms(logits, target, lw, c)
    from egm_unet_b200.loss import loss_terms
    got = loss_terms(logits.cuda(), target.cuda(), lw.cuda(), 255)
    for k in terms:
        assert abs(float(got[k]) - float(terms[k])) < 1e-5 * max(1.0, abs(float(terms[k]))), k
    # CE-only mode (dice=False)
    l2 = EG.criterion({"out": logits.cuda()}, target.cuda(), lw.cuda(), num_classes=c, dice=False, ignore_index=255)
    assert abs(float(l2) - float(terms["ce"])) < 1e-5


def test_eval_metrics_and_sgd():
    import numpy as np, os
    from egm_unet_b200.loss import EvalMetrics
    from egm_unet_b200.abi import call
    from oracle import synth
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_2x64x48.npz"))
    _, target = synth.make_inputs(2, 64, 48, blobs=True)
    m = EvalMetrics(2, 255)
    m.update(torch.from_numpy(fx["logits"]).cuda(), target.cuda())
    assert np.array_equal(m.confusion().cpu().numpy(), fx["mat"])
    assert abs(m.dice - float(fx["dice"])) < 1e-6
    # fused SGD vs torch.optim.SGD
    p = torch.randn(1000); g1, g2 = torch.randn(1000), torch.randn(1000)
    pr = torch.nn.Parameter(p.clone()); opt = torch.optim.SGD([pr], lr=0.02, momentum=0.9, weight_decay=1e-4)
    pc, buf = p.cuda(), torch.zeros(1000, device="cuda")
    hp = torch.tensor([0.02, 0.9, 1e-4, 1.0], device="cuda")
    for g in (g1, g2):
        pr.grad = g.clone(); opt.step()
        call("sgd_step", pc, g.cuda(), buf, 1000, hp)
    assert rel_err(pc.cpu(), pr.detach()) < 1e-6


TC_CASES = [  # cin, cout, k, dil, H, W, N
    (32, 32, 3, 1, 16, 32, 1), (32, 64, 3, 1, 30, 30, 2), (64, 64, 3, 1, 20, 40, 2), (64, 128, 3, 1, 24, 24, 1), (128, 128, 3, 1, 17, 19, 2),
    (256, 256, 3, 1, 15, 15, 1), (512, 256, 3, 1, 12, 12, 1), (256, 128, 3, 1, 16, 16, 1), (64, 32, 3, 1, 33, 47, 1),
    (32, 32, 3, 12, 30, 30, 1), (64, 64, 1, 1, 20, 20, 1), (32, 32, 7, 1, 18, 18, 1), (128, 384, 1, 1, 9, 9, 1),
    (16, 16, 7, 1, 20, 24, 2), (112, 16, 1, 1, 16, 16, 1), (16, 16, 3, 24, 30, 30, 1), (64, 16, 1, 1, 10, 10, 1), (16, 64, 1, 1, 10, 10, 1),
    (256, 512, 3, 1, 10, 10, 1), (224, 32, 1, 1, 12, 12, 1), (48, 48, 3, 1, 18, 20, 1), (80, 96, 3, 1, 12, 20, 2), (448, 64, 1, 1, 9, 9, 1),
    # dilated GRFB branch shapes: thin ones take the one-stage k_conv_tc path and the row-stacked halo wgrad; 64 channels fall back
    (64, 64, 3, 12, 20, 20, 1), (64, 64, 3, 36, 24, 24, 1), (16, 16, 3, 36, 30, 30, 2), (32, 32, 3, 24, 17, 21, 1), (32, 16, 3, 36, 8, 6, 2)]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc(case):
    """tcgen05 implicit-GEMM conv (forward, dgrad, wgrad) vs torch CPU conv on bf16-rounded operands."""
    from egm_unet_b200.engine import conv_module, PackedConv
    from egm_unet_b200 import abi
    cin, cout, k, dil, H, W, N = case
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, bias=False)
    with torch.no_grad():
        m.weight.copy_(m.weight.to(torch.bfloat16).float())
    x = _q(_rand(N, cin, H, W), torch.bfloat16)
    g = _q(_rand(N, cout, H, W, seed=3), torch.bfloat16)
    hs = Harness(torch.bfloat16, use_tc=True)
    xv = hs.var(x)
    mc = m.cuda()
    assert abi.query("conv2d_tc_supported", cin, cout, k, k, dil, 1) == 1
    yv = conv_module(hs.ctx, xv, mc)
    y = hs.out(yv)
    xr = x.clone().requires_grad_(True)
    m = m.cpu()
    yr = m(xr)
    yr.backward(g)
    assert rel_err(y, yr.detach()) < 1e-2, "forward"
    hs.backward(yv, g)
    assert rel_err(hs.grad(xv), xr.grad) < 1e-2, "dgrad"
    assert rel_err(hs.pgrad(mc.weight), m.weight.grad) < 1e-2, "wgrad"


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[2] * c[3] < 40])
def test_conv_lifted_to_tc(case):
    """thin / grouped / odd-channel convs lifted onto the tcgen05 path (zero-padded dense weights): same results as torch."""
    from egm_unet_b200.engine import conv_module
    cin, cout, k, dil, groups, bias, H, W = case
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, groups=groups, bias=bias)
    with torch.no_grad():
        m.weight.copy_(m.weight.to(torch.bfloat16).float())
    x = _q(_rand(2, cin, H, W), torch.bfloat16)
    g = _q(_rand(2, cout, H, W, seed=3), torch.bfloat16)
    hs = Harness(torch.bfloat16, use_tc=True)
    xv = hs.var(x)
    mc = m.cuda()
    yv = conv_module(hs.ctx, xv, mc)
    y = hs.out(yv)
    xr = x.clone().requires_grad_(True)
    m = m.cpu()
    yr = m(xr)
    yr.backward(g)
    assert rel_err(y, yr.detach()) < 1e-2, "forward"
    hs.backward(yv, g)
    assert rel_err(hs.grad(xv), xr.grad) < 1e-2, "dgrad"
    assert rel_err(hs.pgrad(mc.weight), m.weight.grad) < 1e-2, "wgrad"
    if bias:
        assert rel_err(hs.pgrad(mc.bias), m.bias.grad) < 1e-2


@pytest.mark.parametrize("shape", [(2, 80, 80, 32), (3, 37, 41, 16), (1, 120, 120, 64)])
def test_gap_gmp_against_torch(shape):
    """ChannelAttentionModule pooling (src/EGM-UNet.py:1183-1187): global average / max per (n, c) with the FIRST arg-max pixel."""
    from egm_unet_b200 import abi
    n, h, w, c = shape
    torch.manual_seed(5)
    x = torch.randn(n, h, w, c).to(torch.bfloat16)
    x[:, ::7, ::5] = x.amax()                      # exact ties: the first pixel in scan order must win
    xg = x.cuda()
    avg, mx = torch.empty(n * c, device="cuda"), torch.empty(n * c, device="cuda")
    arg = torch.empty(n * c, dtype=torch.int32, device="cuda")
    scratch = torch.empty(n * c * 12 + 16, dtype=torch.uint8, device="cuda")
    abi.call("gap_gmp", xg, avg, mx, arg, scratch, abi.DTYPE_CODE[torch.bfloat16], n, h * w, c)
    xf = x.float().reshape(n, h * w, c)
    assert torch.allclose(avg.cpu().reshape(n, c), xf.mean(1), rtol=1e-4, atol=1e-5)
    assert torch.equal(mx.cpu().reshape(n, c), xf.amax(1))
    first = (xf == xf.amax(1, keepdim=True)).float().argmax(1)          # first index of the maximum
    assert torch.equal(arg.cpu().reshape(n, c).long(), first)


# ------------------------------------------------------------------ BatchNorm fused into the tcgen05 conv epilogue (round 2)
EPI_CASES = [  # cin, cout, k, dil, groups, H, W, N, bias   (train statistics: padded Cout in {16, 32, 64}; every kernel route)
    (32, 32, 3, 1, 1, 37, 29, 2, False),    # halo kernel, ragged tiles
    (64, 32, 3, 1, 1, 48, 40, 2, False), (32, 64, 3, 1, 1, 33, 47, 1, False), (64, 64, 3, 1, 1, 20, 40, 2, False),
    (3, 32, 3, 1, 1, 40, 40, 2, False),     # in_conv.0: lifted 16 -> 32
    (64, 16, 1, 1, 1, 24, 20, 2, False),    # BasicConv 1x1 onto 16
    (64, 8, 1, 1, 1, 24, 20, 2, False),     # lifted thin output (8 valid of 16)
    (8, 8, 1, 1, 1, 19, 23, 2, True),       # edge weight_generator: conv bias + BN
    (16, 16, 3, 12, 1, 30, 30, 2, False),   # dilated GRFB branch: k_conv_tc, nine taps per stage
    (8, 16, 3, 1, 8, 21, 18, 2, False),     # grouped (lifted block-diagonal)
    (128, 64, 3, 1, 1, 24, 24, 1, False),   # k_conv_tc general path (Cin = 128), Cout 64
    (64, 64, 1, 1, 1, 17, 31, 2, False),    # GRFB shortcut 1x1
    (128, 128, 3, 1, 1, 12, 12, 1, False)]  # wide: statistics stay in the streaming kernel (route must still agree)


@pytest.mark.parametrize("case", EPI_CASES)
def test_conv_bn_act_train_stats_in_epilogue(case):
    """conv -> BN(train) -> ReLU with the batch statistics taken in the conv epilogue: output, running statistics, input / weight / gamma / beta
    gradients vs torch CPU; and the same numbers as the unfused route (EGM_NO_BN_FUSE semantics) up to the fp32-vs-bf16 statistics source."""
    from egm_unet_b200.engine import conv_bn_act, ACT_RELU
    cin, cout, k, dil, groups, H, W, N, bias = case
    torch.manual_seed(2)
    conv = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, groups=groups, bias=bias)
    bn = nn.BatchNorm2d(cout, momentum=0.01)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
        bn.weight.copy_(1 + 0.2 * torch.randn(cout)); bn.bias.copy_(0.1 * torch.randn(cout))
    x = _q(_rand(N, cin, H, W), torch.bfloat16)
    g = _q(_rand(N, cout, H, W, seed=3), torch.bfloat16)
    res = {}
    for fuse in (True, False, "fin"):     # "fin": additionally the finalize step inside the apply kernel (egm_bn_finalize_act_fwd)
        hs = Harness(torch.bfloat16, use_tc=True)
        hs.ctx.fuse_bn = bool(fuse)
        hs.ctx.fuse_bn_finalize = fuse == "fin"
        hs.ctx.stats_all = True           # every shape the epilogue supports, not only the ones the profitability rule selects
        cc, bb = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, groups=groups, bias=bias), nn.BatchNorm2d(cout, momentum=0.01)
        cc.load_state_dict(conv.state_dict()); bb.load_state_dict(bn.state_dict())
        cc, bb = cc.cuda(), bb.cuda()
        xv = hs.var(x)
        yv = conv_bn_act(hs.ctx, xv, cc, bb, ACT_RELU)
        y = hs.out(yv)
        hs.backward(yv, g)
        res[fuse] = (y, hs.grad(xv), hs.pgrad(cc.weight), hs.pgrad(bb.weight), hs.pgrad(bb.bias), bb.running_mean.cpu(), bb.running_var.cpu(),
                     int(bb.num_batches_tracked))
    xr = x.clone().requires_grad_(True)
    conv.train(); bn.train()
    yr = torch.relu(bn(conv(xr)))
    yr.backward(g)
    ref = (yr.detach(), xr.grad, conv.weight.grad, bn.weight.grad, bn.bias.grad, bn.running_mean, bn.running_var, 1)
    names = ("y", "dx", "dw", "dgamma", "dbeta", "running_mean", "running_var")
    for nme, a, b in zip(names, res["fin"], res[True]):      # same arithmetic, one launch fewer: identical up to the wgrad atomics' order
        assert float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)) < 1e-3, nme
    assert res["fin"][7] == res[True][7] == 1
    for fuse in (True, False):
        for nme, a, b in zip(names, res[fuse], ref):
            # gradients pass through bf16-stored dy / dz; running statistics: fp32-accumulator statistics (fused) vs statistics of the
            # bf16-rounded z (unfused)
            if nme in ("dx", "dw", "dgamma", "dbeta"):
                # a ReLU mask bit flips wherever bf16 rounding moves z across 0, which changes single gradient elements by O(1):
                # gradients are compared as vectors (relative RMS error and cosine), not element-wise maxima
                rms = float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
                assert rms < 5e-2 and cosine(a, b) > 0.998, (fuse, nme, rms, cosine(a, b))
                continue
            tol = 2e-2 if nme == "y" else (2e-3 if fuse else 6e-3)
            assert rel_err(a, b) < tol, (fuse, nme, rel_err(a, b))
        assert res[fuse][7] == 1
    # statistics from the fp32 accumulators are at least as close to the fp32 reference as statistics of the bf16-rounded z
    assert rel_err(res[True][5], ref[5]) <= rel_err(res[False][5], ref[5]) + 1e-4


@pytest.mark.parametrize("case", EPI_CASES)
@pytest.mark.parametrize("relu", [True, False])
def test_conv_bn_act_eval_folded_into_epilogue(case, relu):
    """inference: BN folded into weights + epilogue bias (+ ReLU) == torch eval conv -> BN -> ReLU, incl. writing into a concat slice"""
    from egm_unet_b200.engine import conv_bn_act, ACT_RELU, ACT_NONE, Ctx, Var
    cin, cout, k, dil, groups, H, W, N, bias = case
    if bias:
        pytest.skip("conv bias + BN only occurs in the sigmoid-gate mode, which is not folded")
    torch.manual_seed(4)
    conv = nn.Conv2d(cin, cout, k, padding=dil * (k - 1) // 2, dilation=dil, groups=groups, bias=False)
    bn = nn.BatchNorm2d(cout)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(cout)); bn.bias.copy_(0.1 * torch.randn(cout))
        bn.running_mean.copy_(0.2 * torch.randn(cout)); bn.running_var.copy_(0.5 + torch.rand(cout))
    x = _q(_rand(N, cin, H, W), torch.bfloat16)
    conv.eval(); bn.eval()
    with torch.no_grad():
        yr = bn(conv(x))
        yr = torch.relu(yr) if relu else yr
    hs = Harness(torch.bfloat16, training=False, use_tc=True)
    hs.ctx.record = False
    cc, bb = conv.cuda(), bn.cuda()
    xv = hs.var(x, needs_grad=False)
    l0 = abi_launches()
    yv = conv_bn_act(hs.ctx, xv, cc, bb, ACT_RELU if relu else ACT_NONE)
    assert rel_err(hs.out(yv), yr) < 2e-2
    # into a slice of a wider (concat) tensor, neighbours untouched
    ctot, off = cout + 24, 8
    cat = Var(torch.full((N, H, W, ctot), 7.0, dtype=torch.bfloat16, device="cuda"))
    conv_bn_act(hs.ctx, xv, cc, bb, ACT_RELU if relu else ACT_NONE, out=cat, out_coff=off)
    got = cat.t.float().cpu().permute(0, 3, 1, 2)
    assert rel_err(got[:, off:off + cout], yr) < 2e-2
    assert float((got[:, :off] - 7).abs().max()) == 0 and float((got[:, off + cout:] - 7).abs().max()) == 0
    assert abi_launches() - l0 < 40


def abi_launches():
    from egm_unet_b200 import abi
    return abi.LAUNCH_COUNTER[0]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,hw", [(64, (70, 61)), (128, (33, 40)), (64, (5, 29)), (256, (30, 30)), (16, (19, 23))])
def test_highpass3_walk_and_gather(C, hw, dtype):
    """out (+)= x - avgpool3x3(x) (EdgeAwareFeatureEnhancer, src/EGM-UNet.py:875,883): the shared-memory row-walking kernel (C % 64 == 0) and the
    gather kernel (thin tensors) against torch, plain and accumulating (the backward form)."""
    from egm_unet_b200 import abi
    from egm_unet_b200.abi import call
    x = _q(_rand(2, C, *hw), dtype)
    ref = x - F.avg_pool2d(x, 3, 1, 1)
    hs = Harness(dtype)
    xv = hs.var(x, needs_grad=False)
    out = torch.empty_like(xv.t)
    n, h, w, c = xv.shape
    call("highpass3", xv.t, out, 0, hs.ctx.code, n, h, w, c)
    got = out.float().cpu().permute(0, 3, 1, 2)
    tol = 1e-5 if dtype == torch.float32 else 1.2e-2
    assert rel_err(got, ref) < tol
    base = _q(_rand(2, C, *hw, seed=9), dtype)
    acc = hs.var(base, needs_grad=False).t.clone()
    call("highpass3", xv.t, acc, 1, hs.ctx.code, n, h, w, c)
    assert rel_err(acc.float().cpu().permute(0, 3, 1, 2), base + ref) < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,K,hw", [(32, 2, (37, 29)), (64, 2, (16, 20)), (32, 4, (9, 11)), (48, 2, (8, 8))])
def test_out_conv_fused_with_layout(C, K, hw, dtype):
    """OutConv (1x1 + bias) reading NHWC and writing fp32 NCHW logits; backward from fp32 NCHW dlogits -> dy, dW, db (csrc/outconv.cu);
    (48, 2) is an unsupported width and must take the generic conv + layout-conversion route with the same results."""
    from egm_unet_b200.engine import out_conv, seed_grad_from_nchw
    torch.manual_seed(3)
    m = nn.Conv2d(C, K, 1)
    x = _q(_rand(2, C, *hw), dtype)
    g = _rand(2, K, *hw, seed=5)
    hs = Harness(dtype)
    xv = hs.var(x)
    mc = nn.Conv2d(C, K, 1)
    mc.load_state_dict(m.state_dict())
    mc = mc.cuda()
    logits, hd = out_conv(hs.ctx, xv, mc)
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (2, K, *hw)
    xr = x.clone().requires_grad_(True)
    yr = m(xr)
    yr.backward(g)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(logits.cpu(), yr.detach()) < tol
    seed_grad_from_nchw(hs.ctx, hd, g.cuda())
    hs.ctx.backward()
    torch.cuda.synchronize()
    assert rel_err(hs.grad(xv), xr.grad) < max(tol, 1e-2 if dtype == torch.bfloat16 else 0)
    assert rel_err(hs.pgrad(mc.weight), m.weight.grad) < tol * 2
    assert rel_err(hs.pgrad(mc.bias), m.bias.grad) < tol * 2
