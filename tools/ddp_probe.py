#!/usr/bin/env python
"""Diagnostic: 2+ ranks, Trainer(use_graph=True) with the NCCL all-reduce captured in the step graph; prints progress per stage and
dumps the Python stacks if a stage stalls.  torchrun --nproc-per-node N tools/ddp_probe.py"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time() % 1000:8.2f}", *a, flush=True)


def main():
    faulthandler.dump_traceback_later(int(os.environ.get("PROBE_STALL", "45")), exit=True)
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import egm_unet_b200 as E
    from egm_unet_b200.trainer import Trainer
    from oracle import synth
    variant = os.environ.get("PROBE_MODEL", "unet")
    model = (E.UNet if variant == "unet" else E.GRFBUNet)(3, 2, base_c=32)
    model.load_state_dict(synth.fill_state_dict(model.state_dict()))
    model = model.to(dev).train()
    if os.environ.get("PROBE_FP32", "0") == "1":
        model.set_check_mode(True)
    tr = Trainer(model, use_graph=True)
    log("trainer built; nccl in graph:", tr._nccl_in_graph)
    size = int(os.environ.get("PROBE_SIZE", "64"))
    for step in range(5):
        image, target = synth.make_inputs(2, size, size, seed=10 * step + rank)
        log("step", step, "issue")
        loss = tr.step(image.to(dev), target.to(dev))
        log("step", step, "issued; graphs:", len(tr._graphs), "comm inside:", [e[4] for e in tr._graphs.values()])
        torch.cuda.synchronize()
        log("step", step, "done loss", float(loss))
    p = tr.store.params.clone()
    lo, hi = p.clone(), p.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    log("replicas identical:", bool(torch.equal(lo, hi)))
    # fingerprint of the trained parameters: equal (up to the split-K atomics' summation order) between runs that differ only in the
    # step's stream concurrency (EGM_WGRAD_STREAM / EGM_BRANCH_PAR_MAXPIX) -- a race in the overlap would show up here
    log("param fingerprint: L2 %.9e  L1 %.9e  dot(arange) %.9e" % (float(p.double().norm()), float(p.double().abs().sum()),
        float((p.double() * torch.arange(p.numel(), device=p.device, dtype=torch.float64).remainder(97.0)).sum())))
    tr.close()
    dist.barrier()
    dist.destroy_process_group()
    log("exit")


if __name__ == "__main__":
    main()
