#!/usr/bin/env python
"""Run ONE eager EGM-UNet train step (batch 16, 480x480, bf16 -- the bench.py workload) inside a cudaProfilerStart/Stop window, after
warm-up steps outside it.  For `ncu --profile-from-start off ...`: the capture then holds exactly the kernels of one step, in order.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/one_step.py
    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_conv_tc|k_wgrad_tc' -o conv python tools/one_step.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--variant", default="egm")
    ap.add_argument("--call-log", default=None, help="write the C-ABI call keys of the profiled step here, in issue order")
    ap.add_argument("--eval", action="store_true", help="inference forward instead of a train step (cfg5 uses --batch 32 --size 1024)")
    a = ap.parse_args()
    import egm_unet_b200 as E
    from egm_unet_b200.trainer import Trainer
    from oracle import synth
    dev = torch.device("cuda")
    cls = {"unet": E.UNet, "egm": E.GRFBUNet, "yuan": E.YuanGRFBUNet}[a.variant]
    model = cls(3, 2, base_c=32)
    model.load_state_dict(synth.fill_state_dict(model.state_dict()))
    model = model.to(dev)
    image, target = synth.make_inputs(a.batch, a.size, a.size)
    image, target = image.to(dev), target.to(dev)
    if a.eval:
        model.eval()
        step = lambda: model(image)["out"]
        ctxm = torch.no_grad()
    else:
        model.train()
        tr = Trainer(model, use_graph=False)
        step = lambda: tr.step(image, target)
        import contextlib
        ctxm = contextlib.nullcontext()
    with ctxm:
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        from egm_unet_b200 import abi
        if a.call_log:
            abi._CALL_LOG = []
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if a.call_log:
            with open(a.call_log, "w") as f:
                f.write("\n".join(abi._CALL_LOG) + "\n")
            abi._CALL_LOG = None
    print("one step done")


if __name__ == "__main__":
    main()
