"""Tape engine: runs the EGM-UNet graph as a sequence of libegm_b200 kernels.

No torch operator runs on the compute path -- torch supplies device memory (caching
allocator), streams and parameter containers only.  Activations are NHWC tensors of the
run dtype (bf16 in production, fp32 in check mode).  Every op appends a backward closure
to the tape; `Tape.backward()` replays them in reverse, writing parameter gradients
straight into the fp32 gradient slots handed out by `Ctx.grad_slot`.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from . import abi
from .abi import call

_SIDE_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}


def _side_streams(device) -> List["torch.cuda.Stream"]:
    """Three side streams per device, shared by every Ctx (so the caching allocator keeps ONE pool per branch slot)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        # high priority like the capture stream: branch kernels ARE the dependency chain while a region is open (the weight-gradient
        # lane, lowest priority, must not get SM slots ahead of them)
        _SIDE_STREAMS[idx] = [torch.cuda.Stream(device=device, priority=-1) for _ in range(3)]
    return _SIDE_STREAMS[idx]


_WGRAD_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}


def _wgrad_streams(device) -> List["torch.cuda.Stream"]:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WGRAD_STREAMS:
        prio = int(os.environ.get("EGM_WGRAD_PRIO", "0"))          # 0 = lowest priority: never ahead of the dgrad chain (A/B knob: -1)
        _WGRAD_STREAMS[idx] = [torch.cuda.Stream(device=device, priority=prio) for _ in range(2)]
    return _WGRAD_STREAMS[idx]


class Parallel:
    """Fork / join of independent branches of the graph onto side streams (the three branches of an EdgeEnhancedGRFB at the small
    resolutions, whose kernels are 3-20 us long and fill a fraction of the 148 SMs).  Forward: `with par.branch(k): ...` runs its body
    on side stream k after the fork point; `par.join()` makes the main stream wait for every branch.  Tape entries pushed inside a
    branch carry (region, k) and `Ctx.backward` replays them the same way, so the branches overlap in both directions -- inside
    the captured CUDA graph they become parallel paths.

    Memory safety under torch's stream-keyed caching allocator: (a) the main stream does nothing between fork and join; (b) a branch
    only touches tensors of its own stream or tensors allocated on the main stream BEFORE the fork, and never writes a tensor another
    branch reads or writes (a gradient shared by several branches is accumulated through per-branch proxy Vars and summed after the
    join, graph.grfb); (c) every later
    use of a side stream starts with a wait on a newer main-stream event, which orders it after whatever the main stream did with
    memory that has meanwhile been returned to that side stream's pool."""

    def __init__(self, ctx: "Ctx", on: bool):
        self.ctx, self.on = ctx, bool(on)
        if self.on:
            ctx._region += 1
            self.rid = ctx._region
            self.main = torch.cuda.current_stream(ctx.device)
            self.fork = torch.cuda.Event()
            self.fork.record(self.main)
            self.used: List["torch.cuda.Stream"] = []

    def branch(self, k: int):
        import contextlib

        @contextlib.contextmanager
        def cm():
            if not self.on:
                yield
                return
            s = _side_streams(self.ctx.device)[k]
            s.wait_event(self.fork)
            prev, self.ctx._cur = self.ctx._cur, (self.rid, k)
            try:
                with torch.cuda.stream(s):
                    yield
            finally:
                self.ctx._cur = prev
            self.used.append(s)
        return cm()

    def join(self):
        if self.on:
            for s in self.used:
                self.main.wait_stream(s)
            self.used = []


ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
MODE_PLAIN, MODE_EDGE_GATE, MODE_RESIDUAL = 0, 1, 2


class Var:
    """A dense NHWC activation [N,H,W,C] and (during backward) its gradient."""
    __slots__ = ("t", "grad", "needs_grad", "gready")

    def __init__(self, t: torch.Tensor, needs_grad: bool = True):
        self.t = t
        self.grad: Optional[torch.Tensor] = None
        self.needs_grad = needs_grad
        self.gready = None            # CUDA event: a side-stream kernel is still writing .grad (Ctx.wgrad_async); waited for on next touch

    def _sync_grad(self):
        if self.gready is not None:
            torch.cuda.current_stream(self.t.device).wait_event(self.gready)
            self.gready = None

    @property
    def shape(self):
        return self.t.shape

    @property
    def C(self) -> int:
        return self.t.shape[3]

    @property
    def M(self) -> int:
        s = self.t.shape
        return s[0] * s[1] * s[2]

    def grad_target(self, partial: bool = False):
        """(tensor, accumulate_flag) to write this Var's gradient into.  partial=True: the writer
        only touches a channel slice, so a fresh buffer is zero-filled."""
        self._sync_grad()
        if self.grad is None:
            self.grad = torch.empty_like(self.t)
            if partial:
                call("memset_zero", self.grad, self.grad.numel() * self.grad.element_size())
                return self.grad, 1
            return self.grad, 0
        return self.grad, 1

    def accum(self, g: torch.Tensor):
        self._sync_grad()
        if self.grad is None:
            self.grad = g
        else:
            call("axpby", self.grad, g, abi.DTYPE_CODE[g.dtype], g.numel(), 1.0, 1.0)


class SkipView:
    """A skip connection of the U that lives in channels [0, c) of its Up level's concat buffer `cat` ([N,H,W,c + c_up]).  The reference
    materialises torch.cat([x2, x1]) in Up.forward (src/EGM-UNet.py:938-947); here the concat is virtual: the DoubleConv's last BN+ReLU
    writes the skip straight into the buffer, MaxPool2d reads (and back-propagates into) the channel-strided view, and Up only adds
    the up-sampled half.  Deliberately NOT a Var: an op that has no strided form fails loudly instead of reading a wrong layout."""
    __slots__ = ("cat", "c")

    def __init__(self, cat: "Var", c: int):
        self.cat, self.c = cat, c

    @property
    def shape(self):
        n, h, w, _ = self.cat.t.shape
        return (n, h, w, self.c)


class Ctx:
    """One forward(/backward) execution."""

    def __init__(self, dtype: torch.dtype, device, training: bool, record: bool, grad_slot: Optional[Callable] = None,
                 use_tc: bool = True):
        self.dtype = dtype
        self.code = abi.DTYPE_CODE[dtype]
        self.device = device
        self.training = training          # BN uses batch statistics and updates running stats
        self.record = record              # build the tape
        self.tape: List[Callable] = []
        self._grad_slot = grad_slot
        self.use_tc = use_tc and dtype == torch.bfloat16
        self.f32 = dict(dtype=torch.float32, device=device)
        self.wplan: Optional["WeightPlan"] = None   # batched weight preparation (Trainer); None = per-conv pack kernels
        # BN finalize inside the apply kernel (egm_bn_finalize_act_fwd).  Opt-in: measured 21.84 vs 21.69 ms/step -- every block redoing
        # the fp64 finalize of its channels costs the small maps more than the 59 saved ~7 us launches give back
        self.fuse_bn_finalize = os.environ.get("EGM_BN_FIN_FUSE", "0") == "1"
        self.fuse_bn = os.environ.get("EGM_NO_BN_FUSE", "0") != "1"   # BN statistics / inference BN+ReLU in the conv epilogue
        # skip connections produced inside the Up concat buffers (SkipView).  Opt-in: measured 23.69 ms/step against 23.56 with the
        # materialised concat (profiles/step_variants_r2.txt) -- what the Up kernel and the gradient slice copy save (0.25 ms) is lost
        # again because BatchNorm on a channel-strided tensor falls off the cp.async.bulk streaming kernels (+0.31 ms at 480^2)
        self.virtual_skip = os.environ.get("EGM_VIRTUAL_SKIP", "0") == "1"
        self.split_upcat = os.environ.get("EGM_NO_SPLIT_UPCAT", "0") != "1"
        self.fuse_edge = os.environ.get("EGM_NO_EDGE_FUSE", "0") != "1"   # edge enhancer: high-pass + 1x1 conv as one composed 3x3 tcgen05 conv
        # epilogue statistics wherever the kernels support them (measured: 23.54 ms/step against 24.04 with the per-layer
        # "profitable" rule, profiles/step_variants_r2.txt); EGM_BN_STATS_PROFITABLE=1 restores the rule
        self.stats_all = os.environ.get("EGM_BN_STATS_PROFITABLE", "0") != "1"
        # branches of a GRFB run on side streams where the maps have at most this many pixels (N*H*W; default: every GRFB level of cfg2); 0 = never
        self.par_maxpix = int(os.environ.get("EGM_BRANCH_PAR_MAXPIX", "921600"))
        # weight-gradient lane: the tcgen05 wgrad kernels of planned convs (nothing reads their packed output before the end of
        # backward) run on one low-priority side stream instead of inside the dgrad -> BN-backward -> dgrad dependency chain
        # (measured 23.13 -> 22.10 ms/step, profiles/step_variants_r2.txt; EGM_WGRAD_STREAM=0 puts them back in line)
        self.wgrad_lane = os.environ.get("EGM_WGRAD_STREAM", "1") != "0"
        self.skip_lane = os.environ.get("EGM_NO_SKIP_LANE", "0") != "1"     # skip-connection gradient slice copies on the lane too
        self.wgrad_lanes = max(1, min(2, int(os.environ.get("EGM_WGRAD_LANES", "1"))))
        self._w_rr = 0
        self._w_hold: list = []           # operands of wgrad kernels still in flight on the lane (kept alive until the join)
        self.grfb_first_in_region = os.environ.get("EGM_GRFB_FIRST_OUTSIDE", "0") != "1"   # first conv of each GRFB branch inside the parallel region
        self._cur = None                  # None = main stream, else (parallel region id, branch) -- the tag of tape entries
        self._region = 0
        self._arenas: Dict[object, list] = {}    # stream slot -> [arena, offset]: a branch zeroes and uses its own arena

    ARENA_DOUBLES = 1 << 16

    def stat_slice(self, n: int) -> torch.Tensor:
        """n zeroed fp64 accumulators for a conv epilogue's BN statistics: slices of ONE arena zeroed by ONE memset per step."""
        if n > self.ARENA_DOUBLES // 8:
            t = self.f64(n)
            call("memset_zero", t, n * 8)
            return t
        slot = None if self._cur is None else self._cur[1]
        a = self._arenas.get(slot)
        if a is None or a[1] + n > self.ARENA_DOUBLES:
            a = self._arenas[slot] = [self.f64(self.ARENA_DOUBLES), 0]
            call("memset_zero", a[0], self.ARENA_DOUBLES * 8)
        t = a[0][a[1]:a[1] + n]
        a[1] += (n + 1) // 2 * 2
        return t

    # ---- allocation helpers
    def empty(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype, device=self.device)

    def zeros_f32(self, n):
        t = torch.empty(n, **self.f32)
        call("memset_zero", t, n * 4)
        return t

    def f64(self, n):
        return torch.empty(n, dtype=torch.float64, device=self.device)

    def grad_slot(self, p: torch.Tensor) -> torch.Tensor:
        """fp32 tensor (same shape as p) that receives dL/dp."""
        return self._grad_slot(p)

    def push(self, fn: Callable):
        if self.record:
            self.tape.append((self._cur, fn))

    def wgrad_async(self, fn: Callable, hold, want_event: bool = False):
        """run fn() (one wgrad launch) on the weight-gradient lane, ordered after everything enqueued on the current stream so far"""
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        w = _wgrad_streams(self.device)[self._w_rr % self.wgrad_lanes]
        self._w_rr += 1
        w.wait_event(ev)
        done = None
        with torch.cuda.stream(w):
            fn()
            if want_event:
                done = torch.cuda.Event()
                done.record(w)
        self._w_hold.append(hold)
        return done

    def join_wgrad(self):
        """make the current stream wait for the weight-gradient lane(s) (before anything reads the packed / bias gradients)"""
        if self._w_hold:
            cur = torch.cuda.current_stream(self.device)
            for w in _wgrad_streams(self.device)[:self.wgrad_lanes]:
                cur.wait_stream(w)
            self._w_hold = []

    def parallel(self, pixels: int) -> Parallel:
        # regions do not nest: the tape replays them as a flat fork / join sequence on the main stream
        return Parallel(self, 0 < pixels <= self.par_maxpix and self._cur is None)

    def backward(self, after_each: Optional[Callable] = None):
        """Replay the tape in reverse.  Entries recorded inside a Parallel region run on their branch's side stream (forked from the
        main stream when the region is entered, joined when it is left); `after_each` (the data-parallel bucket reducer) only runs
        at main-stream points, i.e. when every kernel enqueued so far is ordered before what the main stream does next."""
        main = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        region, fork, used = None, None, {}

        def join():
            for st in used.values():
                main.wait_stream(st)
            used.clear()
        while self.tape:
            tag, fn = self.tape.pop()
            if tag is None:
                if region is not None:
                    join()
                    region = None
                fn()
                if after_each is not None:
                    after_each()
                continue
            rid, k = tag
            if rid != region:
                if region is not None:
                    join()
                region = rid
                fork = torch.cuda.Event()
                fork.record(main)
            if k not in used:
                used[k] = _side_streams(self.device)[k]
                used[k].wait_event(fork)
            prev, self._cur = self._cur, tag
            try:
                with torch.cuda.stream(used[k]):
                    fn()
            finally:
                self._cur = prev
        if region is not None:
            join()
        self.join_wgrad()


def _p(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """parameter tensor -> raw fp32 data (detached)."""
    return None if t is None else t.detach()


# =========================================================================== layout
def from_nchw(ctx: Ctx, x: torch.Tensor) -> Var:
    n, c, h, w = x.shape
    y = ctx.empty(n, h, w, c)
    call("nchw_to_nhwc", x.contiguous(), y, ctx.code, n, c, h, w)
    return Var(y, needs_grad=False)


def to_nchw(ctx: Ctx, x: Var) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty(n, c, h, w, **ctx.f32)
    call("nhwc_to_nchw", x.t, y, ctx.code, n, c, h, w)
    return y


class LogitsHandle:
    """Stands in for the logits Var when OutConv is fused with the layout change (`out_conv`): the backward seed is the fp32 NCHW
    dlogits tensor itself, never rounded to the activation dtype."""
    __slots__ = ("grad_nchw",)

    def __init__(self):
        self.grad_nchw: Optional[torch.Tensor] = None


def out_conv(ctx: Ctx, y: Var, conv: nn.Conv2d):
    """OutConv (src/EGM-UNet.py:952-956): 1x1 conv + bias -> (fp32 NCHW logits, backward-seed handle).  One kernel reads the NHWC
    activation and writes the NCHW logits in fp32; the backward reads fp32 NCHW dlogits and yields dy, dW, db in one pass."""
    n, h, w, c = y.shape
    k = conv.out_channels
    if conv.kernel_size == (1, 1) and conv.groups == 1 and abi.query("outconv_supported", c, k):
        logits = torch.empty(n, k, h, w, **ctx.f32)
        wt, bs = _p(conv.weight), _p(conv.bias)
        call("outconv_fwd", y.t, wt, bs, logits, ctx.code, n, h * w, c, k)
        hd = LogitsHandle()
        if ctx.record:
            def bwd():
                dl, hd.grad_nchw = hd.grad_nchw, None
                if dl is None:
                    return
                gy = ctx.empty(n, h, w, c) if y.needs_grad else None
                call("outconv_bwd", y.t, wt, dl, gy, ctx.grad_slot(conv.weight), ctx.grad_slot(conv.bias) if conv.bias is not None else None,
                     ctx.code, n, h * w, c, k)
                if gy is not None:
                    y.accum(gy)
            ctx.push(bwd)
        return logits, hd
    lv = conv_module(ctx, y, conv)
    return to_nchw(ctx, lv), lv


def seed_grad_from_nchw(ctx: Ctx, x, g_nchw: torch.Tensor):
    if isinstance(x, LogitsHandle):
        x.grad_nchw = g_nchw.contiguous()
        return
    n, h, w, c = x.shape
    g = ctx.empty(n, h, w, c)
    call("nchw_to_nhwc", g_nchw.contiguous(), g, ctx.code, n, c, h, w)
    x.accum(g)


# =========================================================================== convolution
class WSpec:
    """Where a conv's weight comes from: kind 0 = the parameter itself (grouped / thin ones are zero-padded to a dense 16-aligned
    weight), 1 = FusionConv.down on cat[x, x] (W[:, :C] + W[:, C:]), 2 = conv7x7 + conv5x5 + conv3x3 merged into one 7x7."""
    __slots__ = ("kind", "srcs", "bsrcs", "shape", "groups")

    def __init__(self, kind, srcs, bsrcs, shape, groups=1):
        self.kind, self.srcs, self.bsrcs, self.shape, self.groups = kind, tuple(srcs), tuple(b for b in bsrcs if b is not None), tuple(shape), groups

    @property
    def key(self):
        return tuple(id(p) for p in self.srcs)


class WeightJob:
    __slots__ = ("spec", "coutp", "cinp", "wf", "wd", "bpad", "dwp", "gdst")


class WeightPlan:
    """All tcgen05 convs of one model: persistent packed operands + ONE device job table (csrc/wbatch.cu: EgmWJob).
    Step 1 runs the per-conv pack kernels and registers each conv; from step 2 on `prep()` / `unpack()` replace ~450 tiny
    launches by two."""

    def __init__(self):
        self.jobs = {}
        self.ready = False
        self.table = None
        self.bucket_tables = {}        # gradient bucket -> (device job table, n_jobs, unpack elements), see finalize(bucket_of=)
        self.prep_total = self.unpack_total = 0

    def register(self, ctx: "Ctx", spec: WSpec, coutp: int, cinp: int):
        if self.ready or spec.key in self.jobs:
            return
        j = WeightJob()
        j.spec, j.coutp, j.cinp = spec, coutp, cinp
        j.gdst = [ctx.grad_slot(p) for p in spec.srcs]
        self.jobs[spec.key] = j

    def finalize(self, device, bucket_of: Optional[Callable] = None):
        import numpy as np
        assert abi.query("wjob_bytes") == 160
        rows, pb, ub = [], 0, 0
        for j in self.jobs.values():
            co, cig, kh, kw = j.spec.shape
            taps = kh * kw
            n = taps * j.coutp * j.cinp
            j.wf = torch.empty(n, dtype=torch.bfloat16, device=device)
            j.wd = torch.empty(n, dtype=torch.bfloat16, device=device)
            j.dwp = torch.empty(n, dtype=torch.float32, device=device)
            need_b = len(j.spec.bsrcs) > 0 and (j.coutp != co or len(j.spec.bsrcs) > 1)
            j.bpad = torch.empty(j.coutp, dtype=torch.float32, device=device) if need_b else None
            src = [p.detach().data_ptr() for p in j.spec.srcs] + [0, 0]
            bsrc = ([p.detach().data_ptr() for p in j.spec.bsrcs] if need_b else []) + [0, 0, 0]
            g = [t.data_ptr() for t in j.gdst] + [0, 0]
            pack = lambda lo, hi: (lo & 0xffffffff) | (hi << 32)
            rows.append(src[:3] + bsrc[:3] + [j.wf.data_ptr(), j.wd.data_ptr(), j.bpad.data_ptr() if need_b else 0, j.dwp.data_ptr()] + g[:3]
                        + [pack(j.spec.kind, co), pack(cig, j.spec.groups), pack(kh, kw), pack(j.coutp, j.cinp), pb, ub, 0])
            pb += n
            ub += co * cig * (1 if j.spec.kind == 3 else taps)
        self.prep_total, self.unpack_total = pb, ub
        if rows:
            self.table = torch.from_numpy(np.array(rows, dtype=np.uint64).view(np.int64)).to(device)
            if bucket_of is not None:
                # data-parallel overlap: one unpack table per gradient bucket (rows re-based to their own element space).  A job goes
                # with the EARLIEST-launched bucket any of its parameters lives in, so its gradients are in place before any of them
                # is all-reduced (buckets launch strictly in index order, ddp.BucketReducer.flush_ready).
                groups = {}
                for row, j in zip(rows, self.jobs.values()):
                    groups.setdefault(min(bucket_of(p) for p in j.spec.srcs), []).append((row, j))
                for b, items in groups.items():
                    sub, u2 = [], 0
                    for row, j in items:
                        co, cig, kh, kw = j.spec.shape
                        r2 = list(row)
                        r2[18] = u2
                        sub.append(r2)
                        u2 += co * cig * (1 if j.spec.kind == 3 else kh * kw)
                    self.bucket_tables[b] = (torch.from_numpy(np.array(sub, dtype=np.uint64).view(np.int64)).to(device), len(sub), u2)
        self.ready = True

    def prep(self):
        if self.table is not None:
            call("weight_prep_batch", self.table, len(self.jobs), self.prep_total)

    def unpack(self):
        if self.table is not None:
            call("wgrad_unpack_batch", self.table, len(self.jobs), self.unpack_total)

    def unpack_bucket(self, b: int):
        """scatter the packed weight gradients whose parameters belong to gradient bucket `b` (no-op for buckets without tcgen05 convs)"""
        t = self.bucket_tables.get(b)
        if t is not None:
            call("wgrad_unpack_batch", t[0], t[1], t[2])


class PackedConv:
    """Per-forward packed weights of one nn.Conv2d (+ optional weight override for folded / merged kernels)."""

    def __init__(self, ctx: Ctx, weight: torch.Tensor, groups: int, dilation: int, tc_ok: bool):
        co, cig, kh, kw = weight.shape
        self.co, self.cig, self.kh, self.kw, self.groups, self.dil = co, cig, kh, kw, groups, dilation
        self.cin = cig * groups
        self.tc = bool(tc_ok and ctx.use_tc and abi.query("conv2d_tc_supported", self.cin, co, kh, kw, dilation, groups)
                       and abi.query("conv2d_tc_supported", co, self.cin, kh, kw, dilation, groups))
        n = weight.numel()
        if self.tc:
            self.wf = torch.empty(n, dtype=torch.bfloat16, device=ctx.device)
            self.wd = torch.empty(n, dtype=torch.bfloat16, device=ctx.device) if ctx.record else None
            call("pack_conv_weight_tc", weight, self.wf, self.wd, co, self.cin, kh, kw)
        else:
            self.wf = torch.empty(n, **ctx.f32)
            self.wd = torch.empty(n, **ctx.f32) if ctx.record else None
            call("pack_conv_weight", weight, self.wf, self.wd, co, cig, kh, kw, groups)


def _conv_run(ctx, pk: PackedConv, x_t, x_cs, x_co, w, bias, y_t, y_cs, y_co, acc, n, h, wd_, cin, cout):
    if pk.tc and x_cs == cin and x_co == 0 and y_cs == cout and y_co == 0 and not acc:
        call("conv2d_tc", x_t, w, bias, y_t, n, h, wd_, cin, cout, pk.kh, pk.kw, pk.dil)
    else:
        assert not pk.tc, "tensor-core conv needs dense operands"
        call("conv2d_direct", x_t, x_cs, x_co, w, bias, y_t, y_cs, y_co, acc, ctx.code, n, h, wd_, cin, cout, pk.kh, pk.kw, pk.dil, pk.groups)


class Epi:
    """Epilogue request of a tcgen05 conv (egm_conv2d_tc_ex): ReLU (inference, BN folded), BN batch statistics (training),
    and / or writing into a channel slice of an existing tensor."""
    __slots__ = ("relu", "stats", "out", "out_coff", "sums")

    def __init__(self, relu=False, stats=False, out=None, out_coff=0):
        self.relu, self.stats, self.out, self.out_coff, self.sums = relu, stats, out, out_coff, None


def tc_route(ctx: Ctx, cin: int, co: int, kh: int, kw: int, dilation: int, groups: int, sliced: bool, tc_ok: bool = True):
    """(route, padded Cin, padded Cout): 'native' = dense tcgen05 conv, 'lifted' = zero-padded / block-diagonal 16-aligned one, None = CUDA cores"""
    if ctx.use_tc and tc_ok and not sliced and abi.query("conv2d_tc_supported", cin, co, kh, kw, dilation, groups):
        return "native", cin, co
    if ctx.use_tc and tc_ok and abi.query("conv2d_tc_supported", _pad16(cin), _pad16(co), kh, kw, dilation, 1):
        return "lifted", _pad16(cin), _pad16(co)
    return None, cin, co


def conv2d(ctx: Ctx, x: Var, weight: torch.Tensor, bias: Optional[torch.Tensor], *, groups: int = 1, dilation: int = 1,
           x_coff: int = 0, x_cin: Optional[int] = None, wgrad_sink: Optional[Callable] = None,
           bgrad_sink: Optional[Callable] = None, tc_ok: bool = True, wspec: Optional[WSpec] = None, epi: Optional[Epi] = None) -> Var:
    """y = conv2d(x[..., x_coff:x_coff+Cin], weight) + bias  (stride 1, "same" padding).  `weight` is the reference
    [Cout, Cin/groups, kh, kw] fp32 parameter (or a derived tensor; then `wgrad_sink(dw)` receives its gradient).
    `epi` (tcgen05 routes only; the caller checks `tc_route`) asks for a fused epilogue."""
    n, h, w, ctot = x.shape
    wparam, bparam = weight, bias          # gradient slots are keyed by the nn.Parameter objects
    weight, bias = _p(weight), _p(bias)
    co, cig, kh, kw = wspec.shape if weight is None else weight.shape
    cin = cig * groups
    assert (x_cin or ctot - x_coff) == cin, (x.shape, (co, cig, kh, kw), x_coff)
    sliced = not (x_coff == 0 and cin == ctot)
    route, _, _ = tc_route(ctx, cin, co, kh, kw, dilation, groups, sliced, tc_ok)
    native_tc, lifted_tc = route == "native", route == "lifted"
    assert epi is None or route is not None, "fused conv epilogues exist on the tcgen05 routes only"
    # derived weights without a WSpec (ConvTranspose2d repack) keep the per-conv path: their tensors are rebuilt every step
    plan = ctx.wplan if (native_tc or lifted_tc) and ctx.record and (wspec is not None or wgrad_sink is None) else None
    if plan is not None:
        if wspec is None:
            wspec = WSpec(0, (wparam,), (bparam,), (co, cig, kh, kw), groups)
        if plan.ready:
            return _conv2d_planned(ctx, x, plan.jobs[wspec.key], bias, bparam, dilation, x_coff, cin, bgrad_sink, epi)
        plan.register(ctx, wspec, co if native_tc else _pad16(co), cin if native_tc else _pad16(cin))
    assert weight is not None, "a derived weight may only be omitted once the weight plan is ready"
    if lifted_tc:
        return _conv2d_lifted(ctx, x, wparam, bparam, weight, bias, groups, dilation, x_coff, cin, wgrad_sink, bgrad_sink, epi)
    pk = PackedConv(ctx, weight, groups, dilation, tc_ok and not sliced)
    if epi is not None:
        assert pk.tc
        y, ycs, yco = (ctx.empty(n, h, w, co), co, 0) if epi.out is None else (epi.out.t, epi.out.C, epi.out_coff)
        if epi.stats:
            epi.sums = ctx.stat_slice(2 * co)
        call("conv2d_tc_ex", x.t, ctot, x_coff, cin, pk.wf, bias, y, ycs, yco, co, 0, n, h, w, cin, co, pk.kh, pk.kw, pk.dil, int(epi.relu), epi.sums)
        out = Var(y) if epi.out is None else epi.out
    else:
        y = ctx.empty(n, h, w, co)
        _conv_run(ctx, pk, x.t, ctot, x_coff, pk.wf, bias, y, co, 0, 0, n, h, w, cin, co)
        out = Var(y)
    if ctx.record:
        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            # weight gradient
            dwp = torch.empty(weight.numel(), **ctx.f32)
            if pk.tc:
                call("conv2d_wgrad_tc", x.t, dy, dwp, n, h, w, cin, co, kh, kw, dilation)
            else:
                call("conv2d_wgrad_direct", x.t, ctot, x_coff, dy, co, 0, dwp, ctx.code, n, h, w, cin, co, kh, kw, dilation, groups)
            unpack = "unpack_conv_wgrad_tc" if pk.tc else "unpack_conv_wgrad"      # the tcgen05 wgrad packs [taps][Cout][Cin]
            if wgrad_sink is not None:
                dw = torch.empty_like(weight)
                call(unpack, dwp, dw, co, cig, kh, kw, 0.0)
                wgrad_sink(dw)
            else:
                call(unpack, dwp, ctx.grad_slot(wparam), co, cig, kh, kw, 0.0)
            if bias is not None:
                gb = torch.empty(co, **ctx.f32) if bgrad_sink is not None else ctx.grad_slot(bparam)
                call("channel_sum", dy, ctx.code, n * h * w, co, co, 0, ctx.f64(2 * co), gb)
                if bgrad_sink is not None:
                    bgrad_sink(gb)
            if x.needs_grad:
                gx, acc = x.grad_target(partial=sliced)
                if pk.tc and not acc:
                    call("conv2d_tc", dy, pk.wd, None, gx, n, h, w, co, cin, kh, kw, dilation)
                elif pk.tc:
                    tmp = torch.empty_like(x.t)
                    call("conv2d_tc", dy, pk.wd, None, tmp, n, h, w, co, cin, kh, kw, dilation)
                    call("axpby", gx, tmp, ctx.code, tmp.numel(), 1.0, 1.0)
                else:
                    call("conv2d_direct", dy, co, 0, pk.wd, None, gx, ctot, x_coff, acc, ctx.code, n, h, w, co, cin, kh, kw, dilation, groups)
        ctx.push(bwd)
    return out


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def _tc_read_view(ctx, t, n, h, w, ctot, coff, c, cp):
    """(tensor, cstride, coff, valid) of a channel slice as a TMA-readable operand of a tcgen05 conv with `cp` padded channels.
    egm_conv2d_tc_view zero-fills the channels >= valid, so no staging copy is needed unless the slice is not 16-byte aligned."""
    if ctot % 8 == 0 and coff % 8 == 0:
        return t, ctot, coff, c
    tp = ctx.empty(n, h, w, cp)
    if cp != c:
        call("memset_zero", tp, tp.numel() * 2)
    call("copy_slice", t, tp, ctx.code, n * h * w, c, ctot, coff, cp, 0, 0)
    return tp, cp, 0, cp


def _conv_tc_fwd_bwd(ctx, x, x_coff, cin, cinp, co, cop, kh, kw, dilation, wf, wd, bp, bias_present, bparam, bgrad_sink, wgrad_to, epi=None,
                     lane=False):
    """Forward + tape entry of one tcgen05 conv on channel-strided views.  wgrad_to(dwp) receives / names the packed fp32
    gradient buffer: it returns the tensor egm_conv2d_wgrad_tc_view writes and is called again (post=True) afterwards."""
    n, h, w, ctot = x.shape
    M = n * h * w
    sliced = not (x_coff == 0 and cin == ctot)
    xt, xcs, xco, xv = _tc_read_view(ctx, x.t, n, h, w, ctot, x_coff, cin, cinp)
    if epi is None:
        y = ctx.empty(n, h, w, co)
        call("conv2d_tc_view", xt, xcs, xco, xv, wf, bp, y, co, 0, co, 0, n, h, w, cinp, cop, kh, kw, dilation)
        out = Var(y)
    else:
        assert not (ctx.record and epi.out is not None), "writing into a slice is an inference-only epilogue"
        y, ycs, yco = (ctx.empty(n, h, w, co), co, 0) if epi.out is None else (epi.out.t, epi.out.C, epi.out_coff)
        if epi.stats:
            epi.sums = ctx.stat_slice(2 * co)
        call("conv2d_tc_ex", xt, xcs, xco, xv, wf, bp, y, ycs, yco, co, 0, n, h, w, cinp, cop, kh, kw, dilation, int(epi.relu), epi.sums)
        out = Var(y) if epi.out is None else epi.out
    if ctx.record:
        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            dyt, dcs, dco, dv = _tc_read_view(ctx, dy, n, h, w, co, 0, co, cop)
            dwp = wgrad_to(False)
            if lane and ctx.wgrad_lane:       # packed gradient is only read by egm_wgrad_unpack_batch: off the critical path
                ctx.wgrad_async(lambda: call("conv2d_wgrad_tc_view", xt, xcs, xco, xv, dyt, dcs, dco, dv, dwp, n, h, w, cinp, cop, kh, kw, dilation),
                                (xt, dyt, dy))
            else:
                call("conv2d_wgrad_tc_view", xt, xcs, xco, xv, dyt, dcs, dco, dv, dwp, n, h, w, cinp, cop, kh, kw, dilation)
            wgrad_to(True)
            if bias_present:
                gb = torch.empty(co, **ctx.f32) if bgrad_sink is not None else ctx.grad_slot(bparam)
                scratch = ctx.f64(2 * co)
                if lane and ctx.wgrad_lane and bgrad_sink is None:      # the bias gradient goes straight into the flat gradient buffer: same lane
                    ctx.wgrad_async(lambda: call("channel_sum", dy, ctx.code, M, co, co, 0, scratch, gb), (dy, scratch))
                else:
                    call("channel_sum", dy, ctx.code, M, co, co, 0, scratch, gb)
                if bgrad_sink is not None:
                    bgrad_sink(gb)
            if x.needs_grad:
                fresh = x.grad is None
                gx, acc = x.grad_target(partial=sliced)
                if fresh:             # first contribution: dgrad lands directly in (the channel slice of) x's gradient
                    call("conv2d_tc_view", dyt, dcs, dco, dv, wd, None, gx, ctot, x_coff, cin, 0, n, h, w, cop, cinp, kh, kw, dilation)
                else:                 # a read-modify-write epilogue is latency-bound (measured 2.4x slower): add in a streaming pass
                    tmp = ctx.empty(n, h, w, cin)
                    call("conv2d_tc_view", dyt, dcs, dco, dv, wd, None, tmp, cin, 0, cin, 0, n, h, w, cop, cinp, kh, kw, dilation)
                    if sliced:
                        call("copy_slice", tmp, gx, ctx.code, M, cin, cin, 0, ctot, x_coff, 1)
                    else:
                        call("axpby", gx, tmp, ctx.code, tmp.numel(), 1.0, 1.0)
        ctx.push(bwd)
    return out


def _conv2d_lifted(ctx, x, wparam, bparam, weight, bias, groups, dilation, x_coff, cin, wgrad_sink, bgrad_sink, epi=None) -> Var:
    """Thin (C < 16 or C % 16 != 0), grouped or channel-sliced conv on the tcgen05 path: the weight is lifted to a dense
    zero-padded (block-diagonal) 16-aligned one; activations are read / written in place through channel-strided views (TMA
    zero-fills the padding channels).  MMA work on the padding is irrelevant -- these layers are bandwidth-bound."""
    co, cig, kh, kw = weight.shape
    taps = kh * kw
    cinp, cop = _pad16(cin), _pad16(co)
    wpd = torch.empty(cop, cinp, kh, kw, **ctx.f32)
    call("conv_weight_lift", weight, wpd, co, cig, groups, taps, cop, cinp, 0)
    bp = None
    if bias is not None:
        bp = ctx.zeros_f32(cop)
        call("copy_slice", bias, bp, abi.F32, 1, co, co, 0, cop, 0, 0)
    wf = torch.empty(wpd.numel(), dtype=torch.bfloat16, device=ctx.device)
    wd = torch.empty(wpd.numel(), dtype=torch.bfloat16, device=ctx.device) if ctx.record else None
    call("pack_conv_weight_tc", wpd, wf, wd, cop, cinp, kh, kw)
    hold = {}

    def wgrad_to(post):
        if not post:
            hold["dwp"] = torch.empty(wpd.numel(), **ctx.f32)
            return hold["dwp"]
        dwd = torch.empty_like(wpd)
        call("unpack_conv_wgrad_tc", hold.pop("dwp"), dwd, cop, cinp, kh, kw, 0.0)
        dw = torch.empty_like(weight) if wgrad_sink is not None else ctx.grad_slot(wparam)
        call("conv_weight_lift", dw, dwd, co, cig, groups, taps, cop, cinp, 1)
        if wgrad_sink is not None:
            wgrad_sink(dw)
    return _conv_tc_fwd_bwd(ctx, x, x_coff, cin, cinp, co, cop, kh, kw, dilation, wf, wd, bp, bias is not None, bparam, bgrad_sink, wgrad_to, epi)


def _conv2d_planned(ctx, x, job: WeightJob, bias, bparam, dilation, x_coff, cin, bgrad_sink, epi=None) -> Var:
    """tcgen05 conv whose packed operands come from the WeightPlan (filled by egm_weight_prep_batch at the start of the step);
    its weight gradient stays packed in job.dwp until egm_wgrad_unpack_batch at the end of backward."""
    co, cig, kh, kw = job.spec.shape
    bp = job.bpad if job.bpad is not None else bias
    def wgrad_to(post):
        if post:      # the packed gradient is complete: mark the parameters (a DDP bucket reducer listens through grad_slot)
            for p in job.spec.srcs:
                ctx.grad_slot(p)
        return job.dwp
    return _conv_tc_fwd_bwd(ctx, x, x_coff, cin, job.cinp, co, job.coutp, kh, kw, dilation, job.wf, job.wd, bp,
                            bparam is not None or bgrad_sink is not None, bparam, bgrad_sink, wgrad_to, epi, lane=True)


def conv_module(ctx: Ctx, x: Var, m: nn.Conv2d, **kw) -> Var:
    d = m.dilation[0]
    assert m.stride == (1, 1) and m.padding[0] == d * (m.kernel_size[0] - 1) // 2, "only stride-1 'same' convolutions"
    return conv2d(ctx, x, m.weight, m.bias, groups=m.groups, dilation=d, **kw)


# =========================================================================== batch norm (+ activation)
def bn_act(ctx: Ctx, z: Var, bn: nn.BatchNorm2d, act: int, mode: int = MODE_PLAIN, aux: Optional[Var] = None, alpha: float = 0.0,
           out: Optional[Var] = None, out_coff: int = 0, sums: Optional[torch.Tensor] = None) -> Var:
    """y = act(BN(z)) (mode PLAIN) | sigmoid(BN(z))*aux + aux (EDGE_GATE) | relu(alpha*aux + BN(z)) (RESIDUAL).
    With `out`, y is written into out[..., out_coff:out_coff+C] (a concat buffer) and `out` is returned."""
    n, h, w, c = z.shape
    M = n * h * w
    gamma, beta = _p(bn.weight), _p(bn.bias)
    scale, shift, mean, rstd = (torch.empty(c, **ctx.f32) for _ in range(4))
    training = ctx.training or bn.running_mean is None
    mom = 0.1 if bn.momentum is None else bn.momentum
    fused_fin = training and M > 0 and sums is not None and ctx.fuse_bn_finalize
    if fused_fin:                 # statistics came out of the producing conv's epilogue: finalize inside the apply kernel below
        pass
    elif training and M > 0 and sums is not None:
        call("bn_finalize", sums, M, gamma, beta, bn.running_mean, bn.running_var, bn.num_batches_tracked, float(mom), float(bn.eps), 1, c,
             scale, shift, mean, rstd)
    elif training and M > 0:      # statistics + finalize (scale/shift/mean/rstd, running stats) in one launch
        call("bn_stats_finalize", z.t, ctx.code, M, c, c, 0, ctx.f64(2 * c + 1), gamma, beta, bn.running_mean, bn.running_var,
             bn.num_batches_tracked, float(mom), float(bn.eps), scale, shift, mean, rstd)
    else:
        sums = ctx.f64(2 * c)
        if training:
            call("bn_stats", z.t, ctx.code, M, c, c, 0, sums)
        call("bn_finalize", sums, M, gamma, beta, bn.running_mean, bn.running_var, bn.num_batches_tracked if training else None,
             float(mom), float(bn.eps), int(training), c, scale, shift, mean, rstd)
    if out is None:
        y = Var(ctx.empty(n, h, w, c))
        ycs, yco = c, 0
    else:
        y, ycs, yco = out, out.C, out_coff
    if fused_fin:
        call("bn_finalize_act_fwd", sums, M, gamma, beta, bn.running_mean, bn.running_var, bn.num_batches_tracked, float(mom), float(bn.eps),
             scale, shift, mean, rstd, z.t, c, 0, act, mode, aux.t if aux is not None else None, float(alpha), y.t, ycs, yco, ctx.code, M, c)
    else:
        call("bn_act_fwd", z.t, c, 0, scale, shift, act, mode, aux.t if aux is not None else None, float(alpha), y.t, ycs, yco, ctx.code, M, c)
    if ctx.record:
        def bwd():
            y._sync_grad()
            dy = y.grad
            if dy is None:
                return
            if out is None:
                y.grad = None
            auxt = aux.t if aux is not None else None
            coef = torch.empty(3 * c, **ctx.f32)
            if training and M > 0:
                call("bn_act_bwd_reduce_finalize", dy, ycs, yco, z.t, scale, shift, mean, rstd, act, mode, auxt, float(alpha), ctx.code, M, c,
                     ctx.f64(2 * c + 1), gamma, coef, ctx.grad_slot(bn.weight), ctx.grad_slot(bn.bias))
            else:
                s2 = ctx.f64(2 * c)
                call("bn_act_bwd_reduce", dy, ycs, yco, z.t, scale, shift, mean, rstd, act, mode, auxt, float(alpha), ctx.code, M, c, s2)
                call("bn_bwd_finalize", s2, M, gamma, rstd, c, coef, ctx.grad_slot(bn.weight), ctx.grad_slot(bn.bias), int(training))
            dz, _ = z.grad_target()
            daux, dacc = (None, 0)
            if aux is not None and aux.needs_grad:
                daux, dacc = aux.grad_target()
            call("bn_act_bwd_apply", dy, ycs, yco, z.t, scale, shift, mean, rstd, coef, act, mode, auxt, float(alpha), dz, daux, dacc,
                 ctx.code, M, c)
        ctx.push(bwd)
    return y


def conv_for_bn(ctx: Ctx, x: Var, conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """The conv half of conv -> BatchNorm: (z, sums).  `sums` are the batch statistics from the conv epilogue (training, where the kernels
    support it) for `bn_act(..., sums=sums)`, else None.  Lets a caller run the conv early / on another stream than the BN that needs
    more inputs (the GRFB shortcut: relu(scale * fusion + BN(conv(x))))."""
    d = conv.dilation[0]
    assert conv.stride == (1, 1) and conv.padding[0] == d * (conv.kernel_size[0] - 1) // 2, "only stride-1 'same' convolutions"
    co, cig, kh, kw = conv.weight.shape
    cin = cig * conv.groups
    training = ctx.training or bn.running_mean is None
    route, cinp, cop = tc_route(ctx, cin, co, kh, kw, d, conv.groups, x.C != cin)
    if (ctx.fuse_bn and route is not None and kh == kw and training and x.M > 0
            and abi.query("conv2d_tc_stats_supported" if ctx.stats_all else "conv2d_tc_stats_profitable", cinp, cop, kh, kw, d)):
        epi = Epi(stats=True)
        z = conv2d(ctx, x, conv.weight, conv.bias, groups=conv.groups, dilation=d, epi=epi)
        return z, epi.sums
    return conv2d(ctx, x, conv.weight, conv.bias, groups=conv.groups, dilation=d), None


def conv_bn_act(ctx: Ctx, x: Var, conv: nn.Conv2d, bn: nn.BatchNorm2d, act: int, mode: int = MODE_PLAIN, aux: Optional[Var] = None,
                alpha: float = 0.0, out: Optional[Var] = None, out_coff: int = 0) -> Var:
    """act(BN(conv(x))), act in {none, relu}: DoubleConv (src/EGM-UNet.py:44-55) and BasicConv (:958-975) with the BatchNorm fused
    into the tcgen05 conv's epilogue wherever the kernels allow:
      inference  -- gamma*rstd folded into the weights, beta - mean*gamma*rstd as the epilogue bias, ReLU in the epilogue: ONE kernel,
                    written straight into `out[..., out_coff:]` when given; no BN kernel runs.
      training   -- the conv epilogue also takes the per-channel sum / sum of squares of its fp32 accumulators (layers with <= 64
                    output channels, i.e. the large maps), so the statistics pass over z disappears; finalize + apply follow."""
    d = conv.dilation[0]
    assert conv.stride == (1, 1) and conv.padding[0] == d * (conv.kernel_size[0] - 1) // 2, "only stride-1 'same' convolutions"
    co, cig, kh, kw = conv.weight.shape
    cin = cig * conv.groups
    training = ctx.training or bn.running_mean is None
    route, cinp, cop = tc_route(ctx, cin, co, kh, kw, d, conv.groups, x.C != cin)
    if ctx.fuse_bn and route is not None and kh == kw:
        if not training and not ctx.record and conv.bias is None and mode == MODE_PLAIN and act in (ACT_NONE, ACT_RELU):
            scale, shift = torch.empty(co, **ctx.f32), torch.empty(co, **ctx.f32)
            call("bn_finalize", None, 1, _p(bn.weight), _p(bn.bias), bn.running_mean, bn.running_var, None, 0.0, float(bn.eps), 0, co, scale, shift, None, None)
            wfold = torch.empty_like(_p(conv.weight))
            call("scale_rows", _p(conv.weight), scale, wfold, co, cig * kh * kw)
            return conv2d(ctx, x, wfold, shift, groups=conv.groups, dilation=d, epi=Epi(relu=(act == ACT_RELU), out=out, out_coff=out_coff))
        if training and x.M > 0 and abi.query("conv2d_tc_stats_supported" if ctx.stats_all else "conv2d_tc_stats_profitable", cinp, cop, kh, kw, d):
            epi = Epi(stats=True)
            z = conv2d(ctx, x, conv.weight, conv.bias, groups=conv.groups, dilation=d, epi=epi)
            return bn_act(ctx, z, bn, act, mode, aux, alpha, out=out, out_coff=out_coff, sums=epi.sums)
    return bn_act(ctx, conv2d(ctx, x, conv.weight, conv.bias, groups=conv.groups, dilation=d), bn, act, mode, aux, alpha, out=out, out_coff=out_coff)


# =========================================================================== pooling / upsampling
def maxpool2(ctx: Ctx, x) -> Var:
    """MaxPool2d(2, 2); `x` may be a SkipView (read / back-propagated through the channel-strided view of the concat buffer)."""
    n, h, w, c = x.shape
    view = isinstance(x, SkipView)
    xt, xcs = (x.cat.t, x.cat.C) if view else (x.t, c)
    y = Var(ctx.empty(n, h // 2, w // 2, c))
    call("maxpool2x2_fwd_view", xt, xcs, 0, y.t, ctx.code, n, h, w, c)
    if ctx.record:
        def bwd():
            dy, y.grad = y.grad, None
            if dy is None:
                return
            if view:      # the Up block's backward ran first and left the gradient of the whole concat buffer in x.cat.grad
                assert x.cat.grad is not None, "SkipView: the concat buffer has no gradient yet"
                call("maxpool2x2_bwd_view", xt, xcs, 0, dy, x.cat.grad, xcs, 0, 1, ctx.code, n, h, w, c)
                return
            if not x.needs_grad:
                return
            gx, acc = x.grad_target()
            call("maxpool2x2_bwd_view", xt, xcs, 0, dy, gx, c, 0, acc, ctx.code, n, h, w, c)
        ctx.push(bwd)
    return y


def upsample_concat(ctx: Ctx, low: Var, skip) -> Var:
    """cat([skip, pad(bilinear_x2(low))], channel)  -- Up.forward of the reference up to the DoubleConv.  With a SkipView the skip
    half is already in place: only the up-sampled channels are written, and the backward needs no slice copy."""
    n, hl, wl, cu = low.shape
    _, h, w, cs = skip.shape
    if isinstance(skip, SkipView):
        out = skip.cat
        assert out.C == cs + cu, (out.shape, cs, cu)
        call("upsample_concat_fwd", None, low.t, out.t, ctx.code, n, hl, wl, h, w, cs, cu)
        if ctx.record:
            def bwd_v():
                d = out.grad          # stays: MaxPool and the skip's producer still add to / read channels [0, cs)
                if d is None:
                    return
                gl = ctx.empty(n, hl, wl, cu)
                call("upsample_concat_bwd_low", d, gl, ctx.code, n, hl, wl, h, w, cs, cu)
                low.accum(gl)
            ctx.push(bwd_v)
        return out
    out = Var(ctx.empty(n, h, w, cs + cu))
    if ctx.split_upcat and cs % 8 == 0 and cu % 8 == 0:
        # two warp-uniform launches instead of one kernel whose warps mix pure-copy lanes (skip half) with interpolating lanes:
        # the slice copy runs at 4.2 TB/s, the combined kernel at 2.3 TB/s (ncu, profiles/launches_r2_summary.txt)
        call("copy_slice", skip.t, out.t, ctx.code, n * h * w, cs, cs, 0, cs + cu, 0, 0)
        call("upsample_concat_fwd", None, low.t, out.t, ctx.code, n, hl, wl, h, w, cs, cu)
    else:
        call("upsample_concat_fwd", skip.t, low.t, out.t, ctx.code, n, hl, wl, h, w, cs, cu)
    if ctx.record:
        def bwd():
            d, out.grad = out.grad, None
            if d is None:
                return
            gs, acc = skip.grad_target()
            if ctx.wgrad_lane and ctx.skip_lane and ctx.wplan is not None and ctx.wplan.ready and ctx._cur is None:
                # the skip gradient is next touched a whole decoder + encoder level later (MaxPool backward of the level below, then
                # the skip's producer): the slice copy leaves the dependency chain for the side lane; Var.gready orders the next touch
                skip.gready = ctx.wgrad_async(lambda: call("copy_slice", d, gs, ctx.code, n * h * w, cs, cs + cu, 0, cs, 0, acc), (d, gs), want_event=True)
            else:
                call("copy_slice", d, gs, ctx.code, n * h * w, cs, cs + cu, 0, cs, 0, acc)
            gl = ctx.empty(n, hl, wl, cu)
            call("upsample_concat_bwd_low", d, gl, ctx.code, n, hl, wl, h, w, cs, cu)
            low.accum(gl)
        ctx.push(bwd)
    return out


def deconv_concat(ctx: Ctx, low: Var, skip: Var, up: nn.ConvTranspose2d) -> Var:
    """cat([skip, pad(ConvTranspose2d(k=2, s=2)(low))]) -- Up.forward with bilinear=False (src/unet.py:35-37,39-49).
    The transposed conv is a per-pixel GEMM (1x1 conv onto 4*Cout channels, tcgen05 path) followed by a pixel shuffle
    written straight into the concat tensor."""
    n, hl, wl, cin = low.shape
    _, h, w, cs = skip.shape
    cu = up.out_channels
    assert up.kernel_size == (2, 2) and up.stride == (2, 2) and up.in_channels == cin
    wt = torch.empty(4 * cu, cin, 1, 1, **ctx.f32)
    call("deconv_weight_pack", _p(up.weight), wt, cin, cu, 0)
    b4 = None
    if up.bias is not None:
        b4 = torch.empty(4 * cu, **ctx.f32)
        for k in range(4):
            call("copy_slice", _p(up.bias), b4, abi.F32, 1, cu, cu, 0, 4 * cu, k * cu, 0)

    def sink_w(dwt):
        call("deconv_weight_pack", ctx.grad_slot(up.weight), dwt, cin, cu, 1)

    def sink_b(g4):
        gb = ctx.grad_slot(up.bias)
        for k in range(4):
            call("copy_slice", g4, gb, abi.F32, 1, cu, 4 * cu, k * cu, cu, 0, 1 if k else 0)
    z = conv2d(ctx, low, wt, b4, wgrad_sink=sink_w, bgrad_sink=sink_b)
    out = Var(ctx.empty(n, h, w, cs + cu))
    call("pixel_shuffle_concat_fwd", skip.t, z.t, out.t, ctx.code, n, hl, wl, h, w, cs, cu)
    if ctx.record:
        def bwd():
            d, out.grad = out.grad, None
            if d is None:
                return
            gs, acc = skip.grad_target()
            call("copy_slice", d, gs, ctx.code, n * h * w, cs, cs + cu, 0, cs, 0, acc)
            dz = ctx.empty(n, hl, wl, 4 * cu)
            call("pixel_shuffle_concat_bwd", d, dz, ctx.code, n, hl, wl, h, w, cs, cu)
            z.accum(dz)
        ctx.push(bwd)
    return out


def copy_into(ctx: Ctx, src: Var, dst: Var, dst_coff: int, src_coff: int = 0, c: Optional[int] = None):
    """dst[..., dst_coff:dst_coff+c] = src[..., src_coff:src_coff+c]; gradient flows back from dst.grad."""
    c = c or src.C
    M = src.M
    call("copy_slice", src.t, dst.t, ctx.code, M, c, src.C, src_coff, dst.C, dst_coff, 0)
    if ctx.record and src.needs_grad:
        def bwd():
            if dst.grad is None:
                return
            g, acc = src.grad_target(partial=(c != src.C))
            call("copy_slice", dst.grad, g, ctx.code, M, c, dst.C, dst_coff, src.C, src_coff, acc)
        ctx.push(bwd)


def slice_channels(ctx: Ctx, src: Var, coff: int, c: int) -> Var:
    n, h, w, _ = src.shape
    out = Var(ctx.empty(n, h, w, c))
    call("copy_slice", src.t, out.t, ctx.code, src.M, c, src.C, coff, c, 0, 0)
    if ctx.record:
        def bwd():
            d, out.grad = out.grad, None
            if d is None:
                return
            g, acc = src.grad_target(partial=True)
            call("copy_slice", d, g, ctx.code, src.M, c, c, 0, src.C, coff, acc)
        ctx.push(bwd)
    return out


def release_grad(ctx: Ctx, v: Var):
    """Drop v.grad once every producer slice has consumed it (concat buffers)."""
    if ctx.record:
        def bwd():
            v.grad = None
        ctx.push(bwd)


# =========================================================================== EdgeAwareFeatureEnhancer
def edge_enhancer(ctx: Ctx, x: Var, m) -> Var:
    """src/EGM-UNet.py:872-886: y = sigmoid(BN(conv1x1(x - avgpool3(x)))) * x + x.

    bf16 / tcgen05 path, C <= 64 (the halo kernel's range -- the large maps): x - avgpool3(x) is a zero-padded depthwise 3x3 filter, so
    high-pass + 1x1 conv run as ONE 3x3 tcgen05 conv with the composed weight (egm_highpass_compose / WeightPlan kind 3): the TMA-staged
    halo tile feeds the MMAs directly, the BN batch statistics come out of the conv epilogue, and the gate is applied in the second
    (unavoidable: it needs the batch statistics) pass.  The high-pass tensor, its backward pass and 4 HBM round trips are gone.
    Wider maps (C >= 128, 120^2 and below) keep the row-walking high-pass kernel + 1x1 conv: there a 3x3 conv costs more than it saves."""
    n, h, w, c = x.shape
    conv, bn = m.weight_generator[0], m.weight_generator[1]
    if ctx.use_tc and ctx.fuse_edge and c % 16 == 0 and c <= 64 and conv.groups == 1 and abi.query("conv2d_tc_supported", c, c, 3, 3, 1, 1):
        spec = WSpec(3, (conv.weight,), (conv.bias,), (c, c, 3, 3))
        planned = ctx.wplan is not None and ctx.wplan.ready and ctx.record and spec.key in ctx.wplan.jobs
        training = ctx.training or bn.running_mean is None
        epi = Epi(stats=True) if (ctx.fuse_bn and training and x.M > 0 and abi.query("conv2d_tc_stats_supported", c, c, 3, 3, 1)) else None
        if planned:
            z = conv2d(ctx, x, None, conv.bias, wspec=spec, epi=epi)
        else:
            w3 = torch.empty(c, c, 3, 3, **ctx.f32)
            call("highpass_compose", _p(conv.weight), w3, c * c, 0, 1)

            def sink(dw3):
                call("highpass_compose", ctx.grad_slot(conv.weight), dw3, c * c, 1, 0)
            z = conv2d(ctx, x, w3, conv.bias, wgrad_sink=sink, wspec=spec, epi=epi)
        return bn_act(ctx, z, bn, ACT_SIGMOID, MODE_EDGE_GATE, aux=x, sums=epi.sums if epi is not None else None)
    e = Var(ctx.empty(n, h, w, c))
    call("highpass3", x.t, e.t, 0, ctx.code, n, h, w, c)
    if ctx.record:
        def bwd():
            d, e.grad = e.grad, None
            if d is None:
                return
            gx, acc = x.grad_target()
            call("highpass3", d, gx, acc, ctx.code, n, h, w, c)
        ctx.push(bwd)
    return conv_bn_act(ctx, e, conv, bn, ACT_SIGMOID, MODE_EDGE_GATE, aux=x)


# =========================================================================== MCALayer
def mca_layer(ctx: Ctx, x: Var, m) -> Var:
    """src/EGM-UNet.py:686-791 (see csrc/mca.cu)."""
    n, h, w, c = x.shape
    L = abi.query("mca_vec_len", n, h, w, c)
    sums = ctx.f64(2 * L)
    call("mca_stats", x.t, ctx.code, n, h, w, c, sums)
    gates, avg, std = (torch.empty(L, **ctx.f32) for _ in range(3))
    gh, gw, gc = m.h_cw, m.w_hc, m.c_hw
    P = [_p(gh.weight), _p(gh.conv.weight), gh.conv.weight.shape[-1], _p(gw.weight), _p(gw.conv.weight), gw.conv.weight.shape[-1],
         _p(gc.weight), _p(gc.conv.weight), gc.conv.weight.shape[-1]]
    call("mca_gates", sums, n, h, w, c, *P, gates, avg, std)
    y = Var(ctx.empty(n, h, w, c))
    idx = torch.empty(n * h * w * c, dtype=torch.uint8, device=ctx.device) if ctx.record else None
    fused = abi.query("mca_fused_supported", c) and os.environ.get("EGM_MCA_V1", "0") != "1"
    if fused:      # one row-walking pass: u, d^2 and the 3x3 windows never leave the SM
        call("mca_fwd", x.t, gates, y.t, idx, ctx.code, n, h, w, c)
    else:
        call("mca_apply", x.t, gates, y.t, idx, ctx.empty(n, h, w, c), ctx.empty(n, h, w, c), ctx.code, n, h, w, c)
    if ctx.record:
        def bwd():
            dy, y.grad = y.grad, None
            if dy is None:
                return
            du, scratch = ctx.empty(n, h, w, c), ctx.empty(n, h, w, c)
            if fused:
                call("mca_bwd", x.t, gates, dy, idx, du, ctx.code, n, h, w, c)
            else:
                call("mca_bwd_du", x.t, gates, dy, idx, scratch, du, ctx.code, n, h, w, c)
            dG = ctx.f64(L)
            call("mca_prod_sums", du, x.t, ctx.code, n, h, w, c, dG)
            ca, cb = torch.empty(L, **ctx.f32), torch.empty(L, **ctx.f32)
            call("mca_gates_bwd", dG, n, h, w, c, gates, avg, std, *P, ca, cb,
                 ctx.grad_slot(gh.weight), ctx.grad_slot(gh.conv.weight), ctx.grad_slot(gw.weight), ctx.grad_slot(gw.conv.weight),
                 ctx.grad_slot(gc.weight), ctx.grad_slot(gc.conv.weight))
            call("mca_bwd_dx", du, x.t, gates, ca, cb, scratch, ctx.code, n, h, w, c)
            x.accum(scratch)
        ctx.push(bwd)
    return y


