#!/usr/bin/env python
"""bench.py -- EGM-UNet train step, 480x480, batch 16/GPU, bf16 storage / fp32 accumulate (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on rank 0.  A "step" is one full train step (forward + criterion + backward + gradient all-reduce + SGD)
on one synthetic batch.  `value` = images/s over all ranks with inputs resident in HBM; `e2e` = the same step fed from
pinned HOST buffers through the public `Trainer.run` API (H2D of image+target and D2H of the loss inside the timed region).

`roofline` (dominant kernel class = the tcgen05 implicit-GEMM convs of the 18 DoubleConv layers): algorithmic FLOPs of
SURVEY.md s8(d) / the summed device time of those launches.  Device times come from ONE instrumented eager step outside the
timed region: a CUDA-event pair around every C-ABI call on the launch stream, taken in windows of 160 calls with the stream
PRIMED by a spin kernel before each window so that the window is queued before the GPU starts it -- the kernels then run back
to back and the event deltas carry no host launch gaps (round 1's un-primed events absorbed ~16 us per call).  `roofline.hbm` gives the memory-bound families (BatchNorm,
MCALayer, edge high-pass, pool, upsample-concat, loss) as algorithmic bytes / device time against the measured and the
nominal HBM peak.  `cpu_baseline` / `--impl reference`: the UNMODIFIED reference (staged in oracle/_ref by build()) timed on
this box's host cores; falls back to the oracle port when oracle/_ref is absent.
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


def doubleconv_layers(base_c=32):
    c = base_c
    return [("in_conv.0", 3, c, 1), ("in_conv.3", c, c, 1), ("down1.1.0", c, 2 * c, 2), ("down1.1.4", 2 * c, 2 * c, 2),
            ("down2.1.0", 2 * c, 4 * c, 4), ("down2.1.4", 4 * c, 4 * c, 4), ("down3.1.0", 4 * c, 8 * c, 8), ("down3.1.4", 8 * c, 8 * c, 8),
            ("down4.1.0", 8 * c, 8 * c, 16), ("down4.1.4", 8 * c, 8 * c, 16), ("up1.conv.0", 16 * c, 8 * c, 8), ("up1.conv.3", 8 * c, 4 * c, 8),
            ("up2.conv.0", 8 * c, 4 * c, 4), ("up2.conv.3", 4 * c, 2 * c, 4), ("up3.conv.0", 4 * c, 2 * c, 2), ("up3.conv.3", 2 * c, c, 2),
            ("up4.conv.0", 2 * c, c, 1), ("up4.conv.3", c, c, 1)]


def doubleconv_flops_per_image(h=H, w=W, base_c=32, train=True):
    """sum 2*H*W*Cout*9*Cin over the 18 DoubleConv layers; x3 for training minus dgrad of in_conv.0."""
    fwd = sum(2.0 * (h // s) * (w // s) * co * 9 * ci for _, ci, co, s in doubleconv_layers(base_c))
    if not train:
        return fwd
    return 3 * fwd - 2.0 * h * w * base_c * 9 * 3


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


# ------------------------------------------------------------------------------------------------ the CPU arm
def cpu_baseline(steps=2, warmup=1, n=2, variant="egm"):
    """The reference's CPU implementation of the path, timed on the host cores: full fp32 train step (forward + criterion +
    backward + SGD(momentum, weight decay)).  kind "reference": the unmodified reference modules from oracle/_ref (staged by
    __graft_entry__.build(); torch.optim.SGD as in train.py:113-118); kind "port": the oracle restatement (oracle/egm_oracle.py)."""
    from oracle import egm_oracle as O, synth, build_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    image, target = synth.make_inputs(n, H, W)
    lw = torch.tensor([1.0, 2.0])
    ts = []
    if build_ref.available():
        kind = "reference"
        _, tae, _ = build_ref.load()
        model = build_ref.build_model(variant)
        model.load_state_dict(synth.fill_state_dict(model.state_dict()))
        model.train()
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.02, momentum=0.9, weight_decay=1e-4)
        import warnings
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                loss = tae.criterion(model(image), target, lw, num_classes=2, ignore_index=255)
            opt.zero_grad()
            loss.backward()
            opt.step()
            float(loss.item())
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
        what = "unmodified reference (oracle/_ref: src/EGM-UNet.py GRFBUNet + train_utils.criterion + torch.optim.SGD)"
    else:
        kind = "port"
        import egm_unet_b200 as E
        model = E.GRFBUNet(3, 2, base_c=32) if variant == "egm" else E.UNet(3, 2, base_c=32)
        sd = synth.fill_state_dict(model.state_dict())
        mom = {}
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.train_step(sd, mom, image, target, variant, 0.02, 0.9, 1e-4, lw)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
        what = "oracle port (oracle/egm_oracle.py; oracle/_ref not staged)"
    sec = sorted(ts)[len(ts) // 2]
    return {"value": n / sec, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"EGM-UNet fp32 train step (fwd+criterion+bwd+SGD), batch {n} @ {H}x{W} of the batch-{BATCH} workload, median of {steps} after "
                      f"{warmup} warm-up, {what}"}, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 2
    base, sec = cpu_baseline(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)), n=n)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"EGM-UNet (src/EGM-UNet.py GRFBUNet(3,2,base_c=32)) train step, batch {n} sample of the batch-{BATCH} 480x480 workload, CPU"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ per-kernel device times
def primed_profile(step_fn, dev, window=160):
    """{entry point:int args -> {"ms", "calls"}} of one eager step, measured in windows of `window` C-ABI calls: for each window
    the step is run once more with the launch stream blocked by a spin kernel right before the window, so that window's launches
    (and their event records) are all queued before the first one starts -- the kernels run back to back and no host launch
    gap leaks into the event deltas (abi.profile_step)."""
    from egm_unet_b200 import abi
    abi._PROFILE_DETAIL_ALL = True
    try:
        return abi.profile_step(step_fn, window=window)
    finally:
        abi._PROFILE_DETAIL_ALL = False


def _ints(key):
    name, _, rest = key.partition(":")
    return name, [int(v) for v in rest.split(",")] if rest else []


def conv_shape(key):
    """(n, h, w, cin, cout, kh, kw, dil) of a conv2d_tc* / conv2d_wgrad_tc* call key (the view forms put stride/offset ints first)."""
    name, iv = _ints(key)
    if name not in ("conv2d_tc", "conv2d_wgrad_tc", "conv2d_tc_view", "conv2d_wgrad_tc_view", "conv2d_tc_ex") or len(iv) < 8:
        return None
    if name == "conv2d_tc_ex":
        return tuple(iv[-9:-1])
    return tuple(iv[-8:])


def _dc_keys(base_c=32):
    """(H, Cin, Cout) of every launch of the 18 DoubleConv layers: forward and wgrad carry (Cin, Cout), dgrad carries (Cout, Cin);
    in_conv.0 (3 input channels) runs zero-padded as 16 -> 32 and has no dgrad."""
    fw, dg = set(), set()
    for _, ci, co, s in doubleconv_layers(base_c):
        cip = 16 if ci < 16 else ci
        fw.add((H // s, cip, co))
        if ci >= 16:
            dg.add((H // s, co, ci))
    return fw, dg


_DC_FW, _DC_DG = _dc_keys()


def is_doubleconv(key, batch=None):
    """exactly the DoubleConv 3x3 launches (fwd / dgrad / wgrad), matched by (H, Cin, Cout, k=3, dilation 1) against the layer table --
    the GRFB's own 3x3 convs never share a (resolution, channel) signature with a DoubleConv layer"""
    s = conv_shape(key)
    if s is None:
        return False
    n_, h_, w_, ci, co, kh, kw, dil = s
    if not (kh == 3 and kw == 3 and dil == 1 and h_ == w_):
        return False
    if key.startswith("conv2d_wgrad"):
        return (h_, ci, co) in _DC_FW
    return (h_, ci, co) in _DC_FW or (h_, ci, co) in _DC_DG


# algorithmic bytes of the memory-bound families, from the (M, C) each call carries; E = M*C elements, b = bytes/element
def hbm_family_bytes(key, b=2):
    name, iv = _ints(key)
    try:
        if name in ("bn_stats", "bn_stats_finalize"):
            return "batchnorm", iv[1] * iv[2] * b                      # read z
        if name == "bn_act_fwd":
            M, C = iv[-2], iv[-1]
            mode = iv[3]
            return "batchnorm", M * C * b * (3 if mode else 2)         # read z (+aux), write y
        if name in ("bn_act_bwd_reduce", "bn_act_bwd_reduce_finalize"):
            M, C = iv[-2], iv[-1]
            mode = iv[3]
            return "batchnorm", M * C * b * (3 if mode else 2)         # read dy, z (+aux)
        if name == "bn_act_bwd_apply":
            M, C = iv[-2], iv[-1]
            mode = iv[3]
            return "batchnorm", M * C * b * (5 if mode else 3)         # read dy, z (+aux), write dz (+daux)
        if name.startswith("mca_"):
            n_, h_, w_, c_ = iv[-4:]
            E = n_ * h_ * w_ * c_
            per = {"mca_stats": 1, "mca_apply": 2 + 0.5, "mca_fwd": 2 + 0.5, "mca_bwd_du": 2 + 0.5, "mca_prod_sums": 2, "mca_bwd_dx": 3, "mca_bwd": 3.5}.get(name)
            return ("mca", E * b * per) if per else (None, 0)
        if name == "highpass3":
            n_, h_, w_, c_ = iv[-4:]
            return "edge_highpass", n_ * h_ * w_ * c_ * b * 2
        if name == "edge_fused":
            n_, h_, w_, c_ = iv[-4:]
            return "edge_highpass", n_ * h_ * w_ * c_ * b * 2
        if name in ("maxpool2x2_fwd", "maxpool2x2_fwd_view"):
            n_, h_, w_, c_ = iv[-4:]
            return "pool", n_ * h_ * w_ * c_ * b * 1.25
        if name in ("maxpool2x2_bwd", "maxpool2x2_bwd_view"):
            n_, h_, w_, c_ = iv[-4:]
            return "pool", n_ * h_ * w_ * c_ * b * 2.25               # read x, dy/4, write dx
        if name == "upsample_concat_fwd":
            n_, hl, wl, h_, w_, cs, cu = iv[-7:]
            return "upsample_concat", (n_ * hl * wl * cu + n_ * h_ * w_ * cs + n_ * h_ * w_ * (cs + cu)) * b
        if name == "upsample_concat_bwd_low":
            n_, hl, wl, h_, w_, cs, cu = iv[-7:]
            return "upsample_concat", (n_ * h_ * w_ * cu + n_ * hl * wl * cu) * b
        if name == "loss_fwd_bwd":
            n_, c_, h_, w_ = iv[0:4]
            return "loss", n_ * h_ * w_ * (4 * c_ + 8 + 4 * c_)
    except (IndexError, ValueError):
        pass
    return None, 0


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
    ap.add_argument("--pdl", action="store_true", help="launch kernels with programmatic dependent launch (A/B switch; measured slower, off by default)")
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
    if args.pdl:
        abi.query("set_launch_overlap", 1)

    model = (E.GRFBUNet if args.variant == "egm" else E.UNet)(3, 2, base_c=32)
    model.load_state_dict(synth.fill_state_dict(model.state_dict()))
    model = model.to(dev).train()
    if args.check_mode:
        model.set_check_mode(True)
    # Trainer broadcasts rank 0's parameters and buffers at construction (DDP semantics)
    tr = Trainer(model, lr=0.02, momentum=0.9, weight_decay=1e-4, class_weight=[1.0, 2.0], ignore_index=255, use_graph=not args.no_graph)
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

    # ---- per-kernel breakdown of ONE eager step (outside the timed region).  Primary: kernel durations from CUPTI activity records
    # (hardware timestamps per kernel, keyed by the C-ABI call that launched it); fallback: windowed, queue-primed CUDA events.
    tr.use_graph = False                    # the per-kernel breakdown needs real launches
    # ... and every kernel ALONE on the GPU: the production step overlaps the GRFB branches and the weight-gradient lane on side streams
    # (engine.Parallel / Ctx.wgrad_async), which stretches the event delta of each overlapped kernel; the roofline inputs are
    # per-kernel durations, so the instrumented step runs with that concurrency off (the timed region above had it on)
    os.environ["EGM_WGRAD_STREAM"], os.environ["EGM_BRANCH_PAR_MAXPIX"] = "0", "0"
    ev_prof = primed_profile(step_resident, dev)
    launches = max(launches, sum(v["calls"] for v in ev_prof.values()))   # kernels inside one replayed graph == C-ABI launches of one eager step
    prof = abi.profile_step_cupti(step_resident)
    if prof is not None:
        prof2 = abi.profile_step_cupti(step_resident) or prof   # second sample: run-to-run spread of the roofline inputs
        timing = ("kernel durations from CUPTI activity records (torch.profiler) of one eager step, keyed by C-ABI call; "
                  "cross-check: windowed queue-primed CUDA events, which add ~5 us of event/launch overhead per call")
    else:
        prof, prof2 = ev_prof, primed_profile(step_resident, dev)
        timing = ("CUDA events per launch on the launch stream, queue primed by a spin kernel per window of 160 calls (CUPTI unavailable); "
                  "instrumented eager step with the side-stream concurrency of the production step switched off, i.e. each kernel alone")

    def dc_ms(p):
        return sum(v["ms"] for k, v in p.items() if is_doubleconv(k, args.batch))

    step_sum = sum(v["ms"] for v in prof.values())
    tc_all = sum(v["ms"] for k, v in prof.items() if conv_shape(k) is not None)
    tc_ms, tc_ms2 = dc_ms(prof), dc_ms(prof2)
    tc_ms_used = min(tc_ms, tc_ms2) if tc_ms2 > 0 else tc_ms
    flops = doubleconv_flops_per_image() * args.batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
    peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs (kernels timed inside a long step)" if peaks
                else "fallback 1.4 PF sustained / 6.5 TB/s (B200_PROFILING.md)")
    # per-layer DoubleConv table (fwd+dgrad share a key when Cin == Cout; reported per launch shape)
    layer_rows = []
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        if is_doubleconv(k, args.batch):
            n_, h_, w_, ci, co, kh, kw, dil = conv_shape(k)
            fl = 2.0 * n_ * h_ * w_ * ci * co * 9 * v["calls"]
            layer_rows.append({"call": k.split(":")[0], "N": n_, "H": h_, "W": w_, "Cin": ci, "Cout": co, "calls": v["calls"], "ms": round(v["ms"], 4),
                               "tflops": round(fl / (v["ms"] * 1e-3) / 1e12, 1), "frac_of_peak": round(fl / (v["ms"] * 1e-3) / 1e12 / peak_tf, 3)})
    # memory-bound families
    fam = {}
    for k, v in prof.items():
        f, nbytes = hbm_family_bytes(k, 4 if args.check_mode else 2)
        if f:
            d = fam.setdefault(f, {"ms": 0.0, "algorithmic_bytes": 0.0, "launches": 0})
            d["ms"] += v["ms"]; d["algorithmic_bytes"] += nbytes * v["calls"]; d["launches"] += v["calls"]
    for f, d in fam.items():
        gbs = d["algorithmic_bytes"] / max(d["ms"], 1e-9) / 1e6
        d.update({"ms": round(d["ms"], 4), "achieved_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak_hbm, 3),
                  "frac_of_8TBs": round(gbs / 8000.0, 3), "share_of_step": round(d["ms"] / max(step_sum, 1e-9), 4)})
    traffic = None
    try:        # DRAM bytes of the same launches from the committed `ncu --set full` capture (profiles/), if one exists for this round
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv_dram_traffic_r2.json")))
        traffic = tj.get("doubleconv_dram_bytes_per_step")
        traffic_info = {"unit": "bytes per step over the family's %d launches (dram__bytes_read.sum + dram__bytes_write.sum)" % tj["doubleconv"]["launches"],
                        "algorithmic_bytes_per_step": tj.get("doubleconv_algorithmic_bytes_per_step"),
                        "source": "profiles/conv_dram_traffic_r2.json (ncu launch list of tools/one_step.py, cached; bench.py never runs under ncu)"}
    except Exception:
        traffic_info = None
    if tc_ms_used > 0:
        ach = flops / (tc_ms_used * 1e-3) / 1e12
        best = max((r["tflops"] for r in layer_rows), default=0.0)
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "kernel": "tcgen05 implicit-GEMM convs of the 18 DoubleConv layers (k_conv_tc / k_conv_tc_halo fwd+dgrad, k_wgrad_tc_halo)",
                "kernel_ms_per_step": tc_ms_used, "kernel_ms_two_samples": [round(tc_ms, 4), round(tc_ms2, 4)],
                "timing": timing, "kernel_ms_per_step_cuda_events": round(dc_ms(ev_prof), 4),
                "share_of_step": tc_ms_used / max(step_sum, 1e-9), "sum_of_kernel_ms_per_step": step_sum, "all_tcgen05_conv_ms_per_step": tc_all,
                "best_layer_tflops": best, "algorithmic_flops_per_step": flops, "peak_source": peak_src, "hbm": fam,
                "hbm_peak_gbs": peak_hbm, "traffic_info": traffic_info}
    else:
        # no tensor-core kernel ran (fp32 check mode): report the CUDA-core conv against the same peak
        dm = sum(v["ms"] for k, v in prof.items() if k.startswith("conv2d"))
        ach = flops / (max(dm, 1e-9) * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "kernel": "conv2d_direct (CUDA cores; tcgen05 path inactive)", "kernel_ms_per_step": dm, "peak_source": peak_src, "hbm": fam}
    if args.profile_json and rank == 0:
        byname = {}
        for k, v in prof.items():
            d = byname.setdefault(k.split(":")[0], {"ms": 0.0, "calls": 0})
            d["ms"] += v["ms"]; d["calls"] += v["calls"]
        json.dump({"ms_per_step": ms, "sum_of_kernel_ms": step_sum, "by_entry_point": dict(sorted(byname.items(), key=lambda kv: -kv[1]["ms"])),
                   "doubleconv_layers": layer_rows, "hbm_families": fam, "kernels": prof}, open(args.profile_json, "w"), indent=1)

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
                "gpu_launches": int(launches), "cuda_graph": not args.no_graph, "launch_overlap_pdl": bool(args.pdl), "roofline": roof, "clocks": sampler.summary()}
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        tr.close()                      # graphs that hold NCCL kernels must go before the communicator does
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
