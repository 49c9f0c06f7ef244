"""`UNet` with the reference constructor/forward/state_dict (src/unet.py:61-96), running on libegm_b200."""
from egm_unet_b200.models import UNet, DoubleConv, Down, Up, OutConv  # noqa: F401
