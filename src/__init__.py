"""Drop-in for the reference's `src` package (src/__init__.py:1-2 imports `UNet` and `GRFBUNet`; the reference's own file is
broken -- it imports a non-existent src/GRFBUNet.py).  `from src import GRFBUNet, UNet` as in train.py:8."""
from .unet import UNet
from .GRFBUNet import GRFBUNet
from .yuanGRFBUNet import GRFBUNet as YuanGRFBUNet  # noqa: F401

__all__ = ["UNet", "GRFBUNet", "YuanGRFBUNet"]
