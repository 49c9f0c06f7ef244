#!/usr/bin/env python
"""bf16 error budget of the EGM-UNet forward (CPU, oracle only -- test infrastructure, runs without a GPU).

north_star asks bf16 logits within 2e-2 of the fp32 reference.  The CUDA path stores every materialised activation in bf16;
this tool models exactly that with the oracle's storage-rounding switch (oracle/egm_oracle.py: STORAGE) and then puts ONE
storage class at a time back to fp32 to show which stored tensors own the error, and what the CUDA path would pay for keeping
that class in fp32 (extra HBM bytes per train step at batch 16, 480x480 -> ms at the measured 6.55 TB/s).

    python tools/bf16_error_budget.py [--size 160] [--batch 2] [--variant egm] [--eval]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import egm_oracle as O, synth  # noqa: E402

CLASSES = ["dc_z", "dc_y", "bc_z", "bc_y", "conv", "edge", "mca", "mix", "grfb", "rga", "up", "input", "weight"]
# elements per image at 480^2 (base_c = 32) of each class, for the cost column (fwd store + fwd/bwd reloads ~ 3 passes x 2 extra bytes)
E0 = 480 * 480


def class_elems():
    l = [E0 * 32, E0 // 4 * 64, E0 // 16 * 128, E0 // 64 * 256, E0 // 256 * 256]           # one map per level
    dc = 2 * l[0] + 2 * l[1] + 2 * l[2] + 2 * l[3] + 2 * l[4] + (l[3] // 2 + l[3] // 2) + (l[2] // 2 + l[2] // 2) + (l[1] // 2 + l[1] // 2) + 2 * l[0]
    grfb_per = [x for x in l[1:]]
    bc = sum(x * (12 * 0.25) for x in grfb_per)            # 12 BasicConvs of ~C/4 channels each
    return {"dc_z": dc, "dc_y": dc, "bc_z": bc, "bc_y": bc, "conv": sum(x * 1.8 for x in grfb_per), "edge": sum(x * 2.4 for x in grfb_per),
            "mca": sum(grfb_per), "mix": sum(x * 0.25 for x in grfb_per), "grfb": sum(2 * x for x in grfb_per), "rga": l[4] * 2,
            "up": l[4] * 4 + l[3] * 2 + l[2] * 2 + l[1] * 2, "input": E0 * 3, "weight": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=160)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--variant", default="egm")
    ap.add_argument("--eval", action="store_true")
    a = ap.parse_args()
    import egm_unet_b200 as E
    cls = {"unet": E.UNet, "egm": E.GRFBUNet, "yuan": E.YuanGRFBUNet}[a.variant]
    sd = synth.fill_state_dict(cls(3, 2, base_c=32).state_dict())
    image, _ = synth.make_inputs(a.batch, a.size, a.size, blobs=True)
    train = not a.eval

    def run(storage, keep=(), f16=()):
        O.STORAGE, O.STORAGE_FP32, O.STORAGE_FP16 = storage, frozenset(keep), frozenset(f16)
        try:
            with torch.no_grad():
                return O.forward(sd, image, a.variant, train)
        finally:
            O.STORAGE, O.STORAGE_FP32, O.STORAGE_FP16 = None, frozenset(), frozenset()

    ref = run(None)
    span = float(ref.max() - ref.min())

    def stats(o):
        agree = float((o.argmax(1) == ref.argmax(1)).float().mean())
        return float((o - ref).norm() / ref.norm()), float((o - ref).abs().max() / ref.abs().max()), agree

    full = run(torch.bfloat16)
    r0 = stats(full)
    elems = class_elems()
    print(f"# {a.variant} {'train' if train else 'eval'} {a.batch}x3x{a.size}x{a.size}; logit span {span:.3f}")
    print(f"{'storage model':34s} {'RMS rel':>9s} {'max/max':>9s} {'argmax agree':>13s} {'extra ms/step @cfg2':>20s}")
    print(f"{'all classes bf16 (CUDA path)':34s} {r0[0]:9.4f} {r0[1]:9.4f} {r0[2]:13.5f} {0.0:20.2f}")
    rows = []
    for c in CLASSES:
        r = stats(run(torch.bfloat16, (c,)))
        ms = elems[c] * 16 * 2 * 3 / 6.55e12 * 1e3
        rows.append((r0[0] - r[0], c, r, ms))
    for gain, c, r, ms in sorted(rows, reverse=True):
        print(f"{'fp32: ' + c:34s} {r[0]:9.4f} {r[1]:9.4f} {r[2]:13.5f} {ms:20.2f}   (RMS gain {gain:+.4f})")
    both = stats(run(torch.bfloat16, ("dc_y", "bc_y")))
    print(f"{'fp32: dc_y+bc_y (= z-only storage)':34s} {both[0]:9.4f} {both[1]:9.4f} {both[2]:13.5f} {'(saves traffic)':>20s}")
    nonop = ("dc_z", "bc_z", "conv", "mix")
    h = stats(run(torch.bfloat16, (), nonop))
    print(f"{'fp16: pre-BN z (non-MMA-operand)':34s} {h[0]:9.4f} {h[1]:9.4f} {h[2]:13.5f} {0.0:20.2f}")
    h = stats(run(torch.bfloat16, ("input",), nonop))
    print(f"{'  + input as bf16 hi+lo pair':34s} {h[0]:9.4f} {h[1]:9.4f} {h[2]:13.5f} {0.0:20.2f}")
    h = stats(run(torch.bfloat16, (), tuple(CLASSES)))
    print(f"{'fp16: every class':34s} {h[0]:9.4f} {h[1]:9.4f} {h[2]:13.5f}")
    allz = stats(run(torch.bfloat16, ("dc_z", "bc_z", "dc_y", "bc_y")))
    print(f"{'fp32: every conv/BN tensor':34s} {allz[0]:9.4f} {allz[1]:9.4f} {allz[2]:13.5f}")


if __name__ == "__main__":
    main()
