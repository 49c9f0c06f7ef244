"""Fused train step: forward + criterion + backward + (gradient all-reduce) + SGD, all on libegm_b200 kernels.

This is the fast path behind `train_one_epoch` semantics (train_utils/train_and_eval.py:43-75 + SGD of train.py:113-118):
no torch autograd graph, gradients land in one flat fp32 bucket that is all-reduced over NCCL (NVLink/NVSwitch) and consumed
by the fused SGD kernel, and the loss stays on the device (no per-step `.item()` sync -- read it lazily).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import abi
from .abi import call
from .engine import Ctx, WeightPlan, seed_grad_from_nchw
from .graph import net_forward
from .models import _store_for
from .ddp import BucketReducer


class Trainer:
    def __init__(self, model, lr: float = 0.02, momentum: float = 0.9, weight_decay: float = 1e-4,
                 class_weight: Optional[Sequence[float]] = (1.0, 2.0), ignore_index: int = 255, dice: bool = True,
                 process_group=None, num_buckets: int = 4, use_graph: bool = False):
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("Trainer needs the model on a CUDA device (B200); there is no CPU path")
        self.store = _store_for(model)
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.ignore_index, self.dice = ignore_index, dice
        self.cw = None if class_weight is None else torch.tensor(list(class_weight), dtype=torch.float32).to(self.dev)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.mom_buf = torch.empty(self.store.total, dtype=torch.float32, device=self.dev)
        call("memset_zero", self.mom_buf, self.store.total * 4)
        # hyper-parameters reach the device through a RING of pinned slots: the H2D copies are asynchronous, so a slot must not be
        # rewritten before its copy has executed (the LR changes every step under the reference's scheduler)
        self._hp_ring = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(16)]
        self._hp_events = [None] * 16
        self._hp_i = 0
        self.hp = torch.empty(4, dtype=torch.float32, device=self.dev)
        self.loss_terms = None
        self.comm_stream = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
        self.num_buckets = num_buckets
        self.use_graph = use_graph
        self._graph = None            # last replayed CUDA graph of forward + loss + backward (+ SGD when single-rank)
        self._graphs = {}             # (shapes, dtype, store) -> (graph, static image, static target, static loss)
        self._wplan, self._wplan_key = None, None
        self.batch_weights = os.environ.get("EGM_NO_WEIGHT_PLAN", "0") != "1"
        self._gkey = None
        self._calls = 0
        self.reducer = None
        self._nccl_in_graph = False
        if self.world > 1:
            self.reducer = BucketReducer(self.store.grads, self.store.ranges, num_buckets, process_group, self.comm_stream)
            self.sync_replicas()
            # graph mode: capture the bucketed NCCL all-reduces inside the step graph (gloo and other host-driven backends cannot be captured)
            self._nccl_in_graph = dist.get_backend(process_group) == "nccl" and os.environ.get("EGM_DDP_GRAPH_NCCL", "1") != "0"
            if self._nccl_in_graph:       # interpreter exit tears the process group down: the graphs must be gone by then (see close())
                import atexit
                import weakref
                ref = weakref.ref(self)
                atexit.register(lambda: ref() is not None and ref().close())
        self._set_hp()

    def close(self):
        """Drop the captured step graphs (and their private memory pools).  REQUIRED before `dist.destroy_process_group()` when the
        NCCL all-reduces live inside the graphs: destroying a communicator that instantiated graphs still reference hangs."""
        self._graph = None
        self._graphs.clear()
        torch.cuda.synchronize(self.dev)

    def sync_replicas(self, src: int = 0):
        """PyTorch-DDP construction semantics: every rank starts from rank `src`'s parameters AND buffers (BN running statistics,
        num_batches_tracked), whatever seed or checkpoint the other ranks were built from."""
        if self.world <= 1:
            return
        dist.broadcast(self.store.params, src, group=self.pg)
        for b in self.model.buffers():
            dist.broadcast(b, src, group=self.pg)

    def _rebind_store(self):
        """The parameters no longer alias the flat store (model.to(), load_state_dict(assign=True), ...): build a new store and
        re-bind EVERYTHING that pointed into the old one -- bucket reducer, weight plan, captured graph, momentum."""
        old = self.store
        st = self.store = _store_for(self.model)
        if st.total == old.total and st.params.device == self.mom_buf.device:
            pass                                   # same layout: the momentum buffer carries over element for element
        else:
            self.mom_buf = torch.empty(st.total, dtype=torch.float32, device=self.dev)
            call("memset_zero", self.mom_buf, st.total * 4)
        self._graph, self._gkey = None, None
        self._graphs = {}
        self._wplan, self._wplan_key = None, None
        if self.world > 1:
            self.reducer = BucketReducer(st.grads, st.ranges, self.num_buckets, self.pg, self.comm_stream)
        return st

    def _set_hp(self):
        i = self._hp_i = (self._hp_i + 1) % len(self._hp_ring)
        if self._hp_events[i] is not None:
            self._hp_events[i].synchronize()       # 16 updates ago: long done unless nothing ever synchronised
        h = self._hp_ring[i]
        h[0], h[1], h[2], h[3] = self.lr, self.momentum, self.wd, 1.0 / self.world
        self.hp.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._hp_events[i] = ev

    def set_lr(self, lr: float):
        if lr != self.lr:
            self.lr = lr
            self._set_hp()

    def set_hyper(self, lr: float, momentum: float, weight_decay: float):
        if (lr, momentum, weight_decay) != (self.lr, self.momentum, self.wd):
            self.lr, self.momentum, self.wd = lr, momentum, weight_decay
            self._set_hp()

    # ------------------------------------------------------------------ one step
    def forward_backward(self, image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        m, st = self.model, self.store
        if not st.valid():
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("Trainer: parameter storage changed during CUDA-graph capture")
            st = self._rebind_store()
        ctx = Ctx(m.compute_dtype, self.dev, True, True, st.grad_slot, use_tc=m.use_tensor_cores)
        # batched weight preparation: registered during the first (eager) step, two launches per step afterwards.  With a bucket
        # reducer the packed tcgen05 weight gradients are scattered per gradient bucket (WeightPlan.unpack_bucket) right before that
        # bucket's all-reduce is issued, so the exchange still overlaps the rest of backward.
        plan = None
        if self.batch_weights:
            key = (id(st), m.compute_dtype, m.use_tensor_cores)
            if self._wplan is None or self._wplan_key != key:
                self._wplan, self._wplan_key = WeightPlan(), key
            plan = self._wplan
            if not plan.ready and torch.cuda.is_current_stream_capturing():
                plan = None                          # never build a plan (allocations, H2D table copy) inside a capture
            ctx.wplan = plan
            if plan is not None and plan.ready:
                plan.prep()
        logits, lv = net_forward(ctx, m, image, m.variant)
        n, c, h, w = logits.shape
        ws_bytes = abi.query("loss_workspace_bytes", n, c, h, w)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        out = torch.empty(8, dtype=torch.float32, device=self.dev)
        dl = torch.empty_like(logits)
        call("loss_fwd_bwd", logits, target, self.cw, n, c, h, w, self.ignore_index, int(self.dice), 1.0, out, dl, ws, ws_bytes)
        seed_grad_from_nchw(ctx, lv, dl)
        if self.reducer is not None:          # overlap: buckets are all-reduced on the comm stream as backward completes them
            red = self.reducer
            st.on_grad = red.mark
            if plan is not None and plan.ready:      # the packed gradients of a bucket must be complete (weight-gradient lane) before they are scattered
                red.before_launch = lambda b: (ctx.join_wgrad(), plan.unpack_bucket(b))
            else:
                red.before_launch = None
            try:
                ctx.backward(after_each=red.flush_ready)
                red.finish()
            finally:
                st.on_grad = None
                red.before_launch = None
            if plan is not None and not plan.ready:
                owner = red.owner
                plan.finalize(self.dev, bucket_of=lambda p: owner[st.index[id(p)]])
        else:
            ctx.backward()
            if plan is not None:
                if plan.ready:
                    plan.unpack()
                else:
                    plan.finalize(self.dev)
        self.loss_terms = out
        return out[0]

    def step(self, image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """image [N,3,H,W] fp32 cuda, target [N,H,W] int64 cuda -> loss (0-dim device tensor)."""
        self._calls += 1
        if not self.store.valid():
            self._rebind_store()
            self._calls = 1                      # the first step after a rebind runs eagerly (re-registers the weight plan)
        if self.use_graph and self._calls > 1:
            return self._graph_step(image, target)
        loss = self.forward_backward(image, target)
        st = self.store
        call("sgd_step", st.params, st.grads, self.mom_buf, st.total, self.hp)
        return loss

    def _capture_kw(self):
        """With the weight-gradient lane (EGM_WGRAD_STREAM=1) the step is captured on a HIGH-priority stream: the lane's kernels
        (lowest priority) then only take SM slots the dependency chain leaves free.  Kernel nodes inherit their stream's priority."""
        if os.environ.get("EGM_WGRAD_STREAM", "1") == "0":
            return {}
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        return {"stream": self._cap_stream}

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def _graph_step(self, image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """The ~800 kernel launches of a step are captured once per input shape (all kernels take plain pointers and never sync
        or allocate; torch's allocator serves the capture from a private pool) and replayed with new inputs copied into
        static buffers.  With several ranks the graph holds forward+backward; the flat-bucket all-reduce and the fused
        SGD run right after it on the same stream."""
        key = (tuple(image.shape), tuple(target.shape), self.model.compute_dtype, id(self.store))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 3:                          # e.g. full batches + the short last batch of an epoch; bound the pools
                self._graphs.pop(next(iter(self._graphs)))
            s_img, s_tgt = torch.empty_like(image), torch.empty_like(target)
            s_img.copy_(image)
            s_tgt.copy_(target)
            g, loss, comm_inside = None, None, False
            if self.world > 1 and self._nccl_in_graph:
                # NCCL collectives are capturable: the bucketed all-reduces (issued on the comm stream as backward completes each
                # bucket) and the fused SGD become part of the graph, so the gradient exchange overlaps backward in the replayed
                # step exactly as in the eager one
                try:
                    torch.cuda.synchronize(self.dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, **self._capture_kw()):
                        loss = self.forward_backward(s_img, s_tgt)
                        call("sgd_step", self.store.params, self.store.grads, self.mom_buf, self.store.total, self.hp)
                    comm_inside = True
                except Exception as e:             # fall back to exchanging after the replay (and stop trying)
                    import warnings
                    warnings.warn(f"Trainer: capturing the NCCL all-reduce into the CUDA graph failed ({e!r}); exchanging gradients after the replay")
                    self._nccl_in_graph, g = False, None
                    self.reducer.reset()
                    torch.cuda.synchronize(self.dev)
            if g is None:
                reducer, self.reducer = self.reducer, None          # no side-stream traffic inside the capture
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, **self._capture_kw()):
                        loss = self.forward_backward(s_img, s_tgt)
                        if self.world == 1:
                            call("sgd_step", self.store.params, self.store.grads, self.mom_buf, self.store.total, self.hp)
                finally:
                    self.reducer = reducer
            ent = self._graphs[key] = (g, s_img, s_tgt, loss, comm_inside)
        else:
            ent[1].copy_(image, non_blocking=True)
            ent[2].copy_(target, non_blocking=True)
        self._graph = ent[0]
        ent[0].replay()
        if self.world > 1 and not ent[4]:
            dist.all_reduce(self.store.grads, op=dist.ReduceOp.SUM, group=self.pg)
            call("sgd_step", self.store.params, self.store.grads, self.mom_buf, self.store.total, self.hp)
        return ent[3]

    # ------------------------------------------------------------------ host-fed loop (what train_one_epoch does, pipelined)
    def run(self, host_batches):
        """Train on an iterable of (image, target) HOST tensors (ideally pinned) and return the per-step losses as floats.

        Equivalent to the reference loop `for image, target in loader: image.to(device) ...; loss.item()`
        (train_utils/train_and_eval.py:55-73) but without its two serialisation points: the H2D copy of batch i+1 runs on a
        copy stream while batch i computes (double-buffered device staging), and every loss is read back with an async D2H
        into pinned memory that is only waited for one step later ("lazy metric read-back")."""
        out = []
        for ready in self.run_iter(host_batches):
            out.extend(ready)
        return out

    def run_iter(self, host_batches, pre_step=None, post_step=None):
        """Generator form of `run` (what `train_utils.train_one_epoch` drives): yields once per issued step -- plus once at the
        end -- the list of losses (floats, in step order) whose read-back has completed since the previous yield.
        `pre_step()` runs right before a step is launched (set the LR there), `post_step()` right after."""
        dev = self.dev
        copy_stream = getattr(self, "_copy_stream", None) or torch.cuda.Stream(device=dev)
        self._copy_stream = copy_stream
        compute = torch.cuda.current_stream(dev)
        it = iter(host_batches)
        # device staging buffers and pinned loss slots live on the trainer: allocating them (cudaMalloc / cudaHostAlloc) costs
        # milliseconds and would otherwise be paid by every call
        if not hasattr(self, "_stage"):
            self._stage, self._loss_slots = [None, None], [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(3)]
        bufs, ready, freed = self._stage, [None, None], [None, None]
        free_hosts = list(self._loss_slots)

        def prefetch(slot, batch):
            img_h, tgt_h = batch
            if bufs[slot] is None or bufs[slot][0].shape != img_h.shape or bufs[slot][1].shape != tgt_h.shape:
                bufs[slot] = (torch.empty(img_h.shape, dtype=torch.float32, device=dev), torch.empty(tgt_h.shape, dtype=torch.int64, device=dev))
            if freed[slot] is not None:
                copy_stream.wait_event(freed[slot])          # the step that last read this slot has consumed it
            with torch.cuda.stream(copy_stream):
                bufs[slot][0].copy_(img_h, non_blocking=True)
                bufs[slot][1].copy_(tgt_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            ready[slot] = ev

        pending = []
        nxt = next(it, None)
        if nxt is None:
            return
        prefetch(0, nxt)
        i = 0
        while nxt is not None:
            slot = i & 1
            done = []
            compute.wait_event(ready[slot])
            if pre_step is not None:
                pre_step()
            loss = self.step(bufs[slot][0], bufs[slot][1])
            ev = torch.cuda.Event()
            ev.record(compute)
            freed[slot] = ev
            if post_step is not None:
                post_step()
            if len(pending) >= 2:                            # wait for (and recycle the slot of) the loss issued two steps ago
                h, e = pending.pop(0)
                e.synchronize()
                done.append(float(h[0]))
                free_hosts.append(h)
            host = free_hosts.pop()
            host.copy_(loss.detach().reshape(1), non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record(compute)
            pending.append((host, e2))
            nxt = next(it, None)
            if nxt is not None:
                prefetch((i + 1) & 1, nxt)
            i += 1
            yield done
        done = []
        for h, e in pending:
            e.synchronize()
            done.append(float(h[0]))
        yield done
