"""The "yuan" (original) variant: `GRFBUNet` of src/yuanGRFBUNet.py:1474-1512 (no MCALayer in DoubleConv1)."""
from egm_unet_b200.models import YuanGRFBUNet as GRFBUNet  # noqa: F401
