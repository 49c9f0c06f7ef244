"""Recipe for `oracle/_ref`: the UNMODIFIED reference, staged for the GPU box -- TEST / BENCH INFRASTRUCTURE.

The reference (feiyeha/EGM-Unet) is pure Python with no setup.py, so it cannot be pip-installed; its hot path is the handful of
files below.  `/root/reference` exists only in the authoring container, so `__graft_entry__.build()` calls `stage()` there: the
files are copied byte for byte into `oracle/_ref/` (git-ignored: reference sources never enter the history; NOT gpurun-ignored:
the directory travels to the GPU box with the snapshot, like the built .so).  On the GPU box `bench.py --impl reference` and the
config-parity tests import the real reference from there (`load()`), with the two shims SURVEY.md s8c names:
`thop` stubbed, and the model files loaded by path because `src/__init__.py` is broken upstream.

Nothing on the product path imports this module.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("EGM_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["src/unet.py", "src/EGM-UNet.py", "src/yuanGRFBUNet.py", "train_utils/__init__.py", "train_utils/train_and_eval.py",
         "train_utils/dice_coefficient_loss.py", "train_utils/distributed_utils.py", "LICENSE"]


def stage(verbose: bool = True) -> bool:
    """Copy the reference's hot-path files into oracle/_ref (only where /root/reference exists). Returns True if staged."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(REF_DST)
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    if verbose:
        print(f"oracle/_ref: staged {len(FILES)} reference files from {REF_SRC}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "src", "EGM-UNet.py"))


_CACHE = {}


def load():
    """-> (models: {'unet','egm','yuan'} modules, train_and_eval module, distributed_utils module) of the staged reference."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    sys.modules.setdefault("thop", types.SimpleNamespace(profile=lambda *a, **k: None))
    mods = {}
    import contextlib
    import io
    for name, fn in (("unet", "unet.py"), ("egm", "EGM-UNet.py"), ("yuan", "yuanGRFBUNet.py")):
        spec = importlib.util.spec_from_file_location(f"_egm_ref_{name}", os.path.join(REF_DST, "src", fn))
        m = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(m)
        mods[name] = m
    # the reference's train_utils is a package named like ours: load it under a private name
    def load_pkg_module(modname, relpath, package=None):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_DST, relpath),
                                                      submodule_search_locations=[os.path.join(REF_DST, "train_utils")] if package else None)
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    pkg = types.ModuleType("_egm_ref_train_utils")
    pkg.__path__ = [os.path.join(REF_DST, "train_utils")]
    sys.modules["_egm_ref_train_utils"] = pkg
    dcl = load_pkg_module("_egm_ref_train_utils.dice_coefficient_loss", "train_utils/dice_coefficient_loss.py")
    du = load_pkg_module("_egm_ref_train_utils.distributed_utils", "train_utils/distributed_utils.py")
    # train_and_eval.py does `import train_utils.distributed_utils as utils` and `from .dice_coefficient_loss import ...`
    saved = {k: sys.modules.get(k) for k in ("train_utils", "train_utils.distributed_utils", "train_utils.dice_coefficient_loss")}
    sys.modules["train_utils"], sys.modules["train_utils.distributed_utils"], sys.modules["train_utils.dice_coefficient_loss"] = pkg, du, dcl
    try:
        tae = load_pkg_module("_egm_ref_train_utils.train_and_eval", "train_utils/train_and_eval.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["mods"] = (mods, tae, du)
    return _CACHE["mods"]


def build_model(variant: str, **kw):
    """The reference's own nn.Module: UNet / GRFBUNet (EGM) / yuan GRFBUNet with (in=3, classes=2, base_c=32)."""
    mods, _, _ = load()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "unet":
            return mods["unet"].UNet(in_channels=3, num_classes=2, base_c=32, **kw)
        return mods[variant].GRFBUNet(in_channels=3, num_classes=2, base_c=32)


if __name__ == "__main__":
    stage()
