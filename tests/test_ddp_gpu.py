"""GPU: data-parallel correctness of the fused Trainer (the reference ships no DDP driver, the semantics
matched are PyTorch DDP's: replicas start from rank 0's state, gradients are the MEAN over ranks, BatchNorm stays per-rank).

World size 2, one process per rank (torch.multiprocessing.spawn), both Trainer modes:
    eager  -- bucketed all-reduce overlapped with backward on a side stream (ddp.BucketReducer)
    graph  -- CUDA-graph step (what bench.py / SCALE measure): capture + replay, all-reduce and fused SGD behind the replay
Checked per step against the CPU oracle run on the two shards: all-reduced gradient == mean of the per-shard oracle gradients,
parameters after SGD, per-rank BN buffers, and bit-identical parameters on both ranks.

backend "nccl" needs >= 2 GPUs (run with `gpurun --gpus 2`); backend "gloo" drives the same Trainer code with CUDA tensors of
two processes sharing ONE GPU, so the logic is also covered on a single-GPU box.
"""
import os
import traceback

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, backend, mode, port, errq):
    try:
        import torch.distributed as dist
        import egm_unet_b200 as E
        from egm_unet_b200.trainer import Trainer
        from oracle import egm_oracle as O, synth
        from tests.util import cosine
        ndev = torch.cuda.device_count()
        torch.cuda.set_device(rank % ndev)
        dev = torch.device("cuda", rank % ndev)
        dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                **({"device_id": dev} if backend == "nccl" else {}))
        model = E.UNet(3, 2, base_c=32)
        sd0 = synth.fill_state_dict(model.state_dict())
        sd = {k: v.clone() for k, v in sd0.items()}
        if rank != 0:                       # a replica built from a different seed / checkpoint: construction must fix it
            g = torch.Generator().manual_seed(99 + rank)
            for k, v in sd.items():
                if v.dtype.is_floating_point:
                    sd[k] = v + 0.05 * torch.randn(v.shape, generator=g) if "running_var" not in k else v + 0.05
        model.load_state_dict(sd)
        model = model.to(dev).train().set_check_mode(True)
        tr = Trainer(model, lr=0.02, momentum=0.9, weight_decay=1e-4, class_weight=[1.0, 2.0], ignore_index=255, use_graph=(mode == "graph"))
        got = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        for k, v in sd0.items():            # parameters AND buffers equal rank 0's after construction
            assert torch.equal(got[k].to(v.dtype), v), f"rank {rank}: {k} was not broadcast from rank 0"
        # oracle state: shared parameters, per-rank BN buffers
        osd = [{k: v.clone() for k, v in sd0.items()} for _ in range(world)]
        mom = {}
        names = [k for k, _ in model.named_parameters()]
        lw = torch.tensor([1.0, 2.0])
        for step in range(3):
            batches = [synth.make_inputs(2, 48, 48, seed=500 + 10 * step + r) for r in range(world)]
            # every step is checked from the CUDA path's OWN starting point (parameters + momentum): whole-model fp32 gradients are
            # only reproducible to ~1e-2 (ReLU / max-pool kinks), so free-running replicas of oracle and CUDA drift
            # apart after two updates and the third step would compare gradients at different parameters
            cur = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
            for r in range(world):
                for k in names:
                    osd[r][k] = cur[k].clone()
            st = tr.store
            mom = {k: tr.mom_buf[off:off + p.numel()].view(p.shape).detach().cpu().clone()
                   for (k, p), off in zip(model.named_parameters(), st.offsets)} if step > 0 else {}
            loss = tr.step(batches[rank][0].to(dev), batches[rank][1].to(dev))
            torch.cuda.synchronize()
            grads, losses = [], []
            for r in range(world):
                s = osd[r]
                for k in names:
                    s[k].requires_grad_(True)
                    s[k].grad = None
                upd = {}
                lg = O.forward(s, batches[r][0], "unet", True, upd)
                l = O.criterion(lg, batches[r][1], lw)
                l.backward()
                grads.append({k: s[k].grad.detach().clone() for k in names})
                losses.append(float(l))
                for k in names:
                    s[k].requires_grad_(False)
                for k, v in upd.items():
                    s[k] = v.detach()
            assert abs(float(loss) - losses[rank]) <= 1e-4 * abs(losses[rank]), (step, float(loss), losses[rank])
            mean = {k: sum(g[k] for g in grads) / world for k in names}
            # the flat bucket holds the all-reduced SUM (1/world is folded into the SGD kernel)
            num = da = db = 0.0
            for k, p in model.named_parameters():
                g = (tr.store.grad_slot(p).detach().cpu() / world).double().flatten()
                r_ = mean[k].double().flatten()
                num += float(g @ r_); da += float(g @ g); db += float(r_ @ r_)
                if float(r_.norm()) > 1e-4 * max(float(m.norm()) for m in mean.values()):
                    assert abs(float(g.norm()) - float(r_.norm())) <= 3e-2 * float(r_.norm()), (step, k, float(g.norm()), float(r_.norm()))
            assert num / (da * db) ** 0.5 > 0.9995, (step, num / (da * db) ** 0.5)
            # a single-rank gradient would fail this: the two shards' gradients differ
            single = sum(float((grads[0][k] - grads[1][k]).norm()) for k in names) / sum(float(mean[k].norm()) for k in names)
            assert single > 0.05
            with torch.no_grad():           # oracle SGD with the averaged gradient on the shared parameters
                for k in names:
                    d = mean[k] + 1e-4 * osd[0][k]
                    buf = mom.get(k)
                    buf = d.clone() if buf is None else buf.mul_(0.9).add_(d)
                    mom[k] = buf
                    newp = osd[0][k] - 0.02 * buf
                    for r in range(world):
                        osd[r][k] = newp.clone()
            new = model.state_dict()
            for k in names:
                ref = osd[0][k]
                err = float((new[k].cpu() - ref).abs().max() / (ref.abs().max() + 1e-12))
                assert err < 2e-3, (step, k, err)
            for k in sd0:
                if "running_" in k:        # BN statistics are per-rank (no SyncBN)
                    ref = osd[rank][k]
                    assert float((new[k].cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6, (step, k)
            # replicas stay bit-identical
            p = tr.store.params.clone()
            lo, hi = p.clone(), p.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), f"step {step}: replicas diverged"
        if mode == "graph":
            assert tr._graph is not None, "graph mode never replayed a captured step"
            if backend == "nccl":       # the bucketed all-reduces were captured INSIDE the step graph (overlap in the replayed step)
                assert all(e[4] for e in tr._graphs.values()), "NCCL all-reduce was not captured into the graph"
        tr.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def _run(backend, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = 29700 + (os.getpid() % 200) + (0 if backend == "gloo" else 7) + (0 if mode == "eager" else 13)
    procs = [ctx.Process(target=_worker, args=(r, 2, backend, mode, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    import time
    deadline = time.time() + 240          # both workers together; a collective that hangs must not hold the GPU box
    for p in procs:
        p.join(timeout=max(1.0, deadline - time.time()))
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker timed out")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs) or [p.exitcode for p in procs]


@pytest.mark.parametrize("mode", ["eager", "graph"])
def test_world2_gradients_match_oracle_gloo_one_gpu(mode):
    _run("gloo", mode)


@pytest.mark.parametrize("mode", ["eager", "graph"])
def test_world2_gradients_match_oracle_nccl(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run("nccl", mode)
