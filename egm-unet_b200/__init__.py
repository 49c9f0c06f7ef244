"""egm_unet_b200 -- B200-native EGM-UNet hot path (package directory: `egm-unet_b200/`).

Import as `import egm_unet_b200` (the repo-root shim `egm_unet_b200.py` maps the hyphenated
directory onto an importable package name).
"""
from . import abi  # noqa: F401
from .models import UNet, GRFBUNet, YuanGRFBUNet  # noqa: F401
from .loss import criterion, fused_criterion  # noqa: F401

__all__ = ["abi", "UNet", "GRFBUNet", "YuanGRFBUNet", "criterion", "fused_criterion"]
