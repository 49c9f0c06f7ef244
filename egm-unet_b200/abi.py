"""ctypes binding of libegm_b200.so -- the only door to the CUDA kernels.

Prototypes are parsed from include/egm_b200.h, so the header is the single source of
truth for the C ABI (tests check that every declared symbol is exported).  There is no
CPU fallback: a missing library or a non-sm_100 device raises immediately.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(_ROOT, "include", "egm_b200.h")
LIB_PATH = os.environ.get("EGM_LIB") or os.path.join(_HERE, "libegm_b200.so")   # EGM_LIB: a tuning-variant build of the same library

F32, BF16 = 0, 1
DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16}

_CT = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{symbol: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|long\s+long|int)\s+(egm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CT[" ".join(ret.split())]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = " ".join(a.split()[:-1]).replace("const ", "").replace("unsigned ", "")
                    argtypes.append(_CT[base])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(egm-unet_b200/csrc/build.sh). There is no CPU / PyTorch fallback.")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.fn = {}
        for name, (restype, argtypes) in self.protos.items():
            f = getattr(self.cdll, name)          # AttributeError if the .so lacks a declared symbol
            f.restype, f.argtypes = restype, argtypes
            self.fn[name] = f
        v = self.fn["egm_abi_version"]()
        if v != 1:
            raise RuntimeError(f"libegm_b200 ABI version {v} != 1")
        self.device_checked = False

    def last_error(self) -> str:
        return (self.fn["egm_last_error"]() or b"").decode()


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def _conv_arg(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


def call(name: str, *args):
    """Call an int-returning entry point on the current CUDA stream; raise RuntimeError on failure."""
    L = lib()
    if not L.device_checked:
        if not torch.cuda.is_available():
            raise RuntimeError("egm_b200: no CUDA device; this package has no CPU path")
        if L.fn["egm_device_check"]() != 0:
            raise RuntimeError("egm_b200: " + L.last_error())
        L.device_checked = True
    # launch on the current stream of the device that OWNS the operands (not of whatever device happens to be current):
    # `--device cuda:1` without torch.cuda.set_device must not launch into cuda:0's context
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor):
            dev = a.device
            break
    if dev is not None and dev.type == "cuda" and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return call(name, *args)
    if _CUPTI_RANGES:
        key = "egm:" + name + ":" + ",".join(str(a) for a in args if isinstance(a, int))
        with torch.profiler.record_function(key):
            rc = L.fn["egm_" + name](*[_conv_arg(a) for a in args], torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"egm_{name} failed ({rc}): {L.last_error()}")
        LAUNCH_COUNTER[0] += 1
        return
    if _CALL_LOG is not None:
        _CALL_LOG.append(name + ":" + ",".join(str(a) for a in args if isinstance(a, int)))
    stream = torch.cuda.current_stream().cuda_stream
    prof = _PROFILE is not None
    if prof and _WINDOW is not None:
        idx = _WINDOW[2]
        _WINDOW[2] = idx + 1
        if idx == _WINDOW[0]:
            torch.cuda._sleep(_WINDOW[3])      # block the stream while this window's launches are queued behind it
        prof = _WINDOW[0] <= idx < _WINDOW[1]
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.fn["egm_" + name](*[_conv_arg(a) for a in args], stream)
    if rc != 0:
        raise RuntimeError(f"egm_{name} failed ({rc}): {L.last_error()}")
    if prof:
        e1.record()
        key = name
        if _PROFILE_DETAIL_ALL:
            key = name + ":" + ",".join(str(a) for a in args if isinstance(a, int))
        elif _PROFILE_DETAIL and (name.startswith("conv2d") or name.startswith("bn_act") or name in ("bn_stats", "copy_slice", "mca_bwd_du", "highpass3")):
            key = name + ":" + ",".join(str(a) for a in args if isinstance(a, int))
        _PROFILE.append((key, e0, e1))
    LAUNCH_COUNTER[0] += 1


def query(name: str, *args):
    """Call a host-only helper (no stream argument, returns its value)."""
    return lib().fn["egm_" + name](*args)


_CALL_LOG = None      # list: every C-ABI call key in issue order (tools/one_step.py --call-log; matched against an ncu launch list)
_PROFILE = None
_PROFILE_DETAIL = bool(os.environ.get("EGM_PROFILE_DETAIL"))
_PROFILE_DETAIL_ALL = False          # bench.py: every key carries the call's integer arguments (shapes)


_WINDOW = None      # [first call index, end call index, running call index, spin cycles] while a windowed profile runs
_CUPTI_RANGES = False


def profile_step_cupti(fn):
    """Per-C-ABI-call KERNEL durations of one eager fn() from CUPTI activity records (torch.profiler / kineto): every call runs
    inside a `record_function("egm:<entry point>:<int args>")` range and the durations of the kernels it launched (hardware
    timestamps, no event-record or launch-gap overhead) are summed per key.  Returns {key: {"ms", "calls", "kernels"}} or None when
    the profiler recorded no device activity (CUPTI unavailable).  Diagnostic only -- never inside a timed region."""
    global _CUPTI_RANGES
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    _CUPTI_RANGES = True
    try:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
    except Exception:
        return None
    finally:
        _CUPTI_RANGES = False

    def kernel_us(ev, depth=0):
        t = sum(k.duration for k in ev.kernels)
        n = len(ev.kernels)
        if depth < 6:
            for c in ev.cpu_children:
                a, b = kernel_us(c, depth + 1)
                t += a
                n += b
        return t, n
    out = {}
    for ev in prof.events():
        if ev.name.startswith("egm:"):
            us, nk = kernel_us(ev)
            d = out.setdefault(ev.name[4:], {"ms": 0.0, "calls": 0, "kernels": 0})
            d["ms"] += us * 1e-3
            d["calls"] += 1
            d["kernels"] += nk
    if not out or sum(v["kernels"] for v in out.values()) == 0:
        return None
    return dict(sorted(out.items(), key=lambda kv: -kv[1]["ms"]))


def profile_step(fn, window: int = 0, spin_ms: float = 15.0):
    """Run fn() with a CUDA-event pair around every C-ABI call (on the launch stream); returns
    {entry point: {"ms": summed device time, "calls": n}}.  Diagnostic only -- never inside a timed region.

    window = 0: one pass, every call bracketed.  The Python launch loop is slower than the GPU, so the stream runs dry between
    kernels and each event delta also contains that kernel's launch latency (~10-16 us).
    window = W > 0: fn() is run ceil(calls / W) times; pass k brackets only calls [kW, (k+1)W) and blocks the stream with a
    `spin_ms` spin kernel right before them, so those W launches (and their event records) are all queued before the first
    one starts: the kernels run back to back and the deltas are pure device time."""
    global _PROFILE, _WINDOW
    torch.cuda.synchronize()
    recs = []
    try:
        if window <= 0:
            _PROFILE = recs
            fn()
        else:
            cycles = int(spin_ms * 1e-3 * 1.9e9)
            k, total = 0, None
            while total is None or k * window < total:
                _PROFILE = recs
                _WINDOW = [k * window, (k + 1) * window, 0, cycles]
                fn()
                torch.cuda.synchronize()
                total = _WINDOW[2]
                k += 1
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in recs:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["calls"] += 1
    finally:
        _PROFILE = None
        _WINDOW = None
    return dict(sorted(out.items(), key=lambda kv: -kv[1]["ms"]))


# number of C-ABI compute calls issued (bench.py reports launches per step from this)
LAUNCH_COUNTER = [0]
