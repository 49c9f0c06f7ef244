#!/usr/bin/env python
"""Run the other BASELINE.json configs on one B200 and print timings (not the bench line):
   cfg4  yuanGRFBUNet train step bf16, batch 16, 512x512      cfg5  EGM-UNet eval forward bf16, batch 32, 1024x1024
   plus a parity spot-check of cfg5 at batch 1 against the CPU oracle (argmax agreement)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import egm_unet_b200 as E  # noqa: E402
from egm_unet_b200.trainer import Trainer  # noqa: E402
from oracle import egm_oracle as O, synth  # noqa: E402


def timeit(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda")
    # ---- cfg4
    m = E.YuanGRFBUNet(3, 2, base_c=32)
    m.load_state_dict(synth.fill_state_dict(m.state_dict()))
    m = m.to(dev).train()
    tr = Trainer(m, use_graph=True)
    img, tgt = synth.make_inputs(16, 512, 512)
    img, tgt = img.to(dev), tgt.to(dev)
    ms = timeit(lambda: tr.step(img, tgt))
    print(f"cfg4 yuanGRFBUNet train bf16 N=16 512^2: {ms:.2f} ms/step = {16 / ms * 1e3:.1f} img/s, loss {float(tr.loss_terms[0]):.4f}", flush=True)
    del tr, m, img, tgt
    torch.cuda.empty_cache()
    # ---- cfg5
    m = E.GRFBUNet(3, 2, base_c=32)
    sd = synth.fill_state_dict(m.state_dict())
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    img, _ = synth.make_inputs(32, 1024, 1024)
    img = img.to(dev)
    with torch.no_grad():
        ms = timeit(lambda: m(img)["out"], n=2, warm=1)
        print(f"cfg5 EGM-UNet eval bf16 N=32 1024^2: {ms:.2f} ms/fwd = {32 / ms * 1e3:.1f} img/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        one, _ = synth.make_inputs(1, 1024, 1024, blobs=True)
        out = m(one.to(dev))["out"].cpu()
    t0 = time.time()
    with torch.no_grad():
        ref = O.forward(sd, one, "egm", False)
    margin = (ref[:, 0] - ref[:, 1]).abs()
    sure = margin > 0.05 * float(ref.max() - ref.min())
    agree = (out.argmax(1) == ref.argmax(1))
    print(f"cfg5 parity @1x1024^2 vs oracle ({time.time() - t0:.1f}s CPU): argmax agreement {float(agree.float().mean()):.5f} "
          f"(confident pixels {float(agree[sure].float().mean()):.5f}), logits RMS rel {float((out - ref).norm() / ref.norm()):.4f}", flush=True)


if __name__ == "__main__":
    main()
