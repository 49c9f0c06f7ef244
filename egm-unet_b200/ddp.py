"""Data-parallel gradient exchange: bucketed all-reduce of the flat fp32 gradient buffer, overlapped with backward.

The reference ships no DDP driver; the semantics matched are PyTorch DDP's: every rank runs the whole
step on its own batch (BatchNorm statistics stay per-rank) and parameter gradients are averaged.  Parameters live in ONE
flat buffer in registration order (in_conv ... out_conv); backward produces them in reverse, so buckets are contiguous
slices taken from the END of the buffer.  As soon as every parameter of a bucket has been written, the slice is
all-reduced (NCCL over NVLink/NVSwitch on a side stream, gated by a CUDA event) while backward keeps running; the
1/world_size scaling is folded into the fused SGD kernel.  The same code runs over gloo on CPU tensors for the tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def make_buckets(ranges: Sequence[Tuple[int, int]], num_buckets: int) -> List[Tuple[int, int, List[int]]]:
    """ranges[i] = (offset, padded_numel) of parameter i in registration order.  Returns buckets in LAUNCH order
    (last parameters first) as (start, end, [param indices]), balanced by element count."""
    total = sum(n for _, n in ranges)
    target = max(1, total // max(1, num_buckets))
    buckets, cur, cur_n = [], [], 0
    for i in range(len(ranges) - 1, -1, -1):
        cur.append(i)
        cur_n += ranges[i][1]
        if cur_n >= target and len(buckets) < num_buckets - 1:
            buckets.append(cur)
            cur, cur_n = [], 0
    if cur:
        buckets.append(cur)
    out = []
    for idx in buckets:
        lo = min(ranges[i][0] for i in idx)
        hi = max(ranges[i][0] + ranges[i][1] for i in idx)
        out.append((lo, hi, sorted(idx)))
    return out


class BucketReducer:
    def __init__(self, flat_grads: torch.Tensor, ranges: Sequence[Tuple[int, int]], num_buckets: int = 4, group=None,
                 comm_stream: Optional["torch.cuda.Stream"] = None):
        self.flat, self.group, self.comm_stream = flat_grads, group, comm_stream
        self.buckets = make_buckets(ranges, num_buckets)
        self.owner = {}
        for b, (_, _, idx) in enumerate(self.buckets):
            for i in idx:
                self.owner[i] = b
        self.reset()

    def reset(self):
        self.pending = [set(idx) for _, _, idx in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []
        self.order = []

    def mark(self, param_index: int):
        """the gradient of parameter `param_index` is being written by the kernels enqueued in the current tape closure"""
        self.pending[self.owner[param_index]].discard(param_index)

    def flush_ready(self):
        """call after a tape closure returned (all its kernels are enqueued): launch every bucket that became complete.
        Buckets go out strictly in index order (last parameters first), so `before_launch(b)` hooks may rely on every earlier
        bucket having been handled."""
        for b, (lo, hi, _) in enumerate(self.buckets):
            if self.launched[b]:
                continue
            if self.pending[b]:
                break
            self._launch(b, lo, hi)

    # optional hook(bucket index) run on the compute stream right before a bucket's all-reduce is issued: the Trainer uses it to
    # scatter the packed tcgen05 weight gradients of that bucket into the flat buffer (engine.WeightPlan.unpack_bucket)
    before_launch = None

    def _launch(self, b, lo, hi):
        self.launched[b] = True
        self.order.append(b)
        if self.before_launch is not None:
            self.before_launch(b)
        view = self.flat[lo:hi]
        if self.flat.is_cuda and self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """launch whatever is left, then make the current stream wait for every reduction"""
        for b, (lo, hi, _) in enumerate(self.buckets):
            if not self.launched[b]:
                self._launch(b, lo, hi)
        for w in self.works:
            w.wait()
        if self.flat.is_cuda and self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        order = self.order
        self.reset()
        return order
