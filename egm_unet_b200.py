"""Import shim: makes the package directory `egm-unet_b200/` importable as `egm_unet_b200`."""
import os as _os
import sys as _sys

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "egm-unet_b200")]
__package__ = "egm_unet_b200"
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
_sys.modules.setdefault("egm_unet_b200", _sys.modules[__name__])
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
