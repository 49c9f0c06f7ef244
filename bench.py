#!/usr/bin/env python
"""bench.py -- EGM-UNet train step, 480x480, batch 16/GPU, bf16 storage / fp32 accumulate (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on rank 0.  A "step" is one full train step (forward + criterion + backward + gradient all-reduce + SGD)
on one synthetic batch.  `value` = images/s over all ranks with inputs resident in HBM; `e2e` = the same step fed from
pinned HOST buffers (H2D of image+target and D2H of the loss inside the timed region).  `roofline` is for the dominant
kernel class (tcgen05 DoubleConv convs: algorithmic FLOPs of SURVEY.md s8(d) / their summed CUDA-event time);
`cpu_baseline` is the CPU oracle port timed on this box's host cores.  `--impl reference` times the reference's CPU
path (the oracle port -- the reference is pure Python and cannot be pip-installed or compiled) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 480
BATCH = 16
METRIC = "egm_unet_train_images_per_sec_480_bf16"


def doubleconv_flops_per_image(h=H, w=W, base_c=32, train=True):
    """sum 2*H*W*Cout*9*Cin over the 18 DoubleConv layers; x3 for training minus dgrad of in_conv.0."""
    c = base_c
    layers = [(3, c, 1), (c, c, 1), (c, 2 * c, 2), (2 * c, 2 * c, 2), (2 * c, 4 * c, 4), (4 * c, 4 * c, 4), (4 * c, 8 * c, 8), (8 * c, 8 * c, 8),
              (8 * c, 8 * c, 16), (8 * c, 8 * c, 16), (16 * c, 8 * c, 8), (8 * c, 4 * c, 8), (8 * c, 4 * c, 4), (4 * c, 2 * c, 4),
              (4 * c, 2 * c, 2), (2 * c, c, 2), (2 * c, c, 1), (c, c, 1)]
    fwd = sum(2.0 * (h // s) * (w // s) * co * 9 * ci for ci, co, s in layers)
    if not train:
        return fwd
    return 3 * fwd - 2.0 * h * w * c * 9 * 3


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nme, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(steps=2, warmup=1, n=2, variant="egm"):
    """The oracle port (CPU restatement of the reference, oracle/egm_oracle.py) timed on the host cores: full train step."""
    from oracle import egm_oracle as O, synth
    import egm_unet_b200 as E
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = E.GRFBUNet(3, 2, base_c=32) if variant == "egm" else E.UNet(3, 2, base_c=32)
    sd = synth.fill_state_dict(model.state_dict())
    mom = {}
    image, target = synth.make_inputs(n, H, W)
    lw = torch.tensor([1.0, 2.0])
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(sd, mom, image, target, variant, 0.02, 0.9, 1e-4, lw)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sorted(ts)[len(ts) // 2]
    return {"value": n / sec, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"EGM-UNet fp32 train step (fwd+criterion+bwd+SGD), batch {n} @ {H}x{W}, median of {steps} after {warmup} warm-up, oracle/egm_oracle.py"}, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 2
    base, sec = cpu_baseline(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)), n=n)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"EGM-UNet (src/EGM-UNet.py GRFBUNet(3,2,base_c=32)) train step, batch {n} sample of the batch-16 480x480 workload, CPU"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--variant", default="egm")
    ap.add_argument("--check-mode", action="store_true", help="fp32 check mode (not a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel CUDA-event breakdown of one step here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import egm_unet_b200 as E
    from egm_unet_b200 import abi
    from egm_unet_b200.trainer import Trainer
    from oracle import synth          # synthetic inputs / deterministic weights only (test infrastructure, not the measured path)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps

    model = (E.GRFBUNet if args.variant == "egm" else E.UNet)(3, 2, base_c=32)
    model.load_state_dict(synth.fill_state_dict(model.state_dict()))
    model = model.to(dev).train()
    if args.check_mode:
        model.set_check_mode(True)
    tr = Trainer(model, lr=0.02, momentum=0.9, weight_decay=1e-4, class_weight=[1.0, 2.0], ignore_index=255, use_graph=not args.no_graph)
    tr_eager = tr if args.no_graph else None
    if world > 1:   # identical replicas: broadcast rank 0's parameters / buffers once
        dist.broadcast(tr.store.params, 0)
        for b in model.buffers():
            dist.broadcast(b, 0)
    image_h, target_h = synth.make_inputs(args.batch, H, W, seed=1234 + rank)
    image_h, target_h = image_h.pin_memory(), target_h.pin_memory()
    image, target = image_h.to(dev), target_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def step_resident():
        tr.step(image, target)

    e2e_steps = max(2, K)

    def run_e2e(steps=None):
        # the public host-fed API: every step copies ITS batch from pinned host memory (prefetched on a copy stream while the
        # previous step computes) and its loss is read back to the host (async D2H, waited for one step later)
        steps = steps or e2e_steps
        losses = tr.run((image_h, target_h) for _ in range(steps))
        assert len(losses) == steps and all(v == v for v in losses)

    for _ in range(W_):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = abi.LAUNCH_COUNTER[0]
    ms = timed(step_resident, K)
    launches = (abi.LAUNCH_COUNTER[0] - l0) // max(K, 1)
    run_e2e(min(W_, 3))                     # warm-up of the host-fed path (staging buffers, pinned loss slots, copy stream)
    ms_e2e = timed(run_e2e, 1) / e2e_steps
    sampler.stop_flag = True
    sampler.join(timeout=3)

    # ---- per-kernel breakdown of ONE step with CUDA events on the launch stream (outside the timed region)
    tr.use_graph = False                    # the per-kernel event breakdown needs real launches
    launches_eager0 = abi.LAUNCH_COUNTER[0]
    abi._PROFILE_DETAIL = True              # keys carry the conv shapes so the DoubleConv launches can be told apart
    prof = abi.profile_step(step_resident)
    if launches == 0:
        launches = abi.LAUNCH_COUNTER[0] - launches_eager0     # kernels inside one replayed graph == kernels of one eager step

    def is_doubleconv(key):
        """DoubleConv 3x3 launches (fwd / dgrad / wgrad): k=3, dilation 1, >= 32 channels on both sides (in_conv.0 runs as 16->32)."""
        name, _, shape = key.partition(":")
        if name not in ("conv2d_tc", "conv2d_wgrad_tc", "conv2d_tc_view", "conv2d_wgrad_tc_view") or not shape:
            return False
        n_, h_, w_, ci, co, kh, kw, dil = [int(v) for v in shape.split(",")][-8:]     # the view forms put their stride/offset ints first
        return kh == 3 and dil == 1 and ((min(ci, co) >= 32) or (h_ == H and {ci, co} == {16, 32}))

    step_sum = sum(v["ms"] for v in prof.values())
    tc_all = sum(v["ms"] for k, v in prof.items() if k.startswith("conv2d_tc") or k.startswith("conv2d_wgrad_tc"))
    tc_ms = sum(v["ms"] for k, v in prof.items() if is_doubleconv(k))
    flops = doubleconv_flops_per_image() * args.batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
    if tc_ms > 0:
        ach = flops / (tc_ms * 1e-3) / 1e12
        # best single layer (largest-FLOP launches are compute-bound; the 32/64-channel 480^2 / 240^2 layers are HBM-bound, AI < ridge)
        best = 0.0
        for k, v in prof.items():
            if is_doubleconv(k):
                n_, h_, w_, ci, co = [int(x) for x in k.split(":")[1].split(",")][-8:-3]
                best = max(best, 2.0 * n_ * h_ * w_ * ci * co * 9 * v["calls"] / (v["ms"] * 1e-3) / 1e12)
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "kernel": "tcgen05 implicit-GEMM convs of the 18 DoubleConv layers (k_conv_tc / k_conv_tc_halo fwd+dgrad, k_wgrad_tc_halo)",
                "kernel_ms_per_step": tc_ms, "share_of_step": tc_ms / max(step_sum, 1e-9), "all_tcgen05_conv_ms_per_step": tc_all,
                "best_layer_tflops": best, "algorithmic_flops_per_step": flops, "peak_source": peak_src}
    else:
        # no tensor-core kernel ran (fp32 check mode): report the CUDA-core conv against the same peak
        dm = sum(v["ms"] for k, v in prof.items() if k.startswith("conv2d"))
        ach = flops / (max(dm, 1e-9) * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "kernel": "conv2d_direct (CUDA cores; tcgen05 path inactive)", "kernel_ms_per_step": dm, "peak_source": peak_src}
    if args.profile_json and rank == 0:
        json.dump({"ms_per_step": ms, "kernels": prof}, open(args.profile_json, "w"), indent=1)

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline()
        gb = args.batch * world
        h2d = image_h.numel() * 4 + target_h.numel() * 8
        line = {"metric": METRIC, "value": gb / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.check_mode else "bf16", "data": "synthetic",
                "config": {"workload": f"EGM-UNet (GRFBUNet(3,2,base_c=32)) train step: fwd + criterion(CE+Dice+laplace+lap+sobel) + bwd + SGD, "
                                       f"batch {args.batch}/GPU, 3x{H}x{W}, 2 classes", "global_batch": gb, "parallelism": f"dp{world}",
                           "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush needed"},
                "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
                "gpu_launches": int(launches), "cuda_graph": not args.no_graph, "roofline": roof, "clocks": sampler.summary()}
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
