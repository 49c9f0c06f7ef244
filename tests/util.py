"""Shared helpers for the GPU parity tests: drive engine ops from NCHW CPU tensors and read results/gradients back."""
import torch

import egm_unet_b200  # noqa: F401
from egm_unet_b200 import abi
from egm_unet_b200.abi import call
from egm_unet_b200.engine import Ctx, Var, from_nchw, to_nchw, seed_grad_from_nchw


class Harness:
    def __init__(self, dtype=torch.float32, training=True, use_tc=True):
        self.slots = {}
        self.ctx = Ctx(dtype, torch.device("cuda"), training, True, self.slot, use_tc=use_tc)

    def slot(self, p):
        if id(p) not in self.slots:
            t = torch.empty(p.shape, dtype=torch.float32, device="cuda")
            call("memset_zero", t, max(t.numel(), 1) * 4)
            self.slots[id(p)] = t
        return self.slots[id(p)]

    def var(self, x_nchw, needs_grad=True) -> Var:
        v = from_nchw(self.ctx, x_nchw.cuda().float())
        v.needs_grad = needs_grad
        return v

    def out(self, v: Var):
        return to_nchw(self.ctx, v).cpu()

    def backward(self, v: Var, g_nchw):
        seed_grad_from_nchw(self.ctx, v, g_nchw.cuda().float())
        self.ctx.backward()
        torch.cuda.synchronize()

    def grad(self, v: Var):
        n, h, w, c = v.shape
        y = torch.empty(n, c, h, w, dtype=torch.float32, device="cuda")
        call("nhwc_to_nchw", v.grad, y, self.ctx.code, n, c, h, w)
        return y.cpu()

    def pgrad(self, p):
        return self.slots[id(p)].cpu()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))


TOL = {torch.float32: 2e-4, torch.bfloat16: 4e-2}
