"""EGM-UNet: `GRFBUNet` of the reference's src/EGM-UNet.py:1503-1541, running on libegm_b200."""
from egm_unet_b200.models import GRFBUNet  # noqa: F401
