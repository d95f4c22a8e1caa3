"""Drop-in `utils` package (reference: utils/__init__.py:1-3 re-exports LBSWsmpl.compute_lbswField,
FindSurfacePs.* and utils.*): the ray/surface finder, Jacobian / cardinal-ray / normal helpers, small
math, and the driver-facing config / checkpoint helpers."""
from . import FindSurfacePs as _find
from . import utils as _math
from .LBSWsmpl import compute_lbswField

for _mod in (_find, _math):
    for _name in dir(_mod):
        if not _name.startswith("_"):
            globals()[_name] = getattr(_mod, _name)
del _mod, _name
