"""Reference names (MCAcc/interp2x_boundary3d.py:8,22) for the upsample + boundary op in _autograd.py."""
from ._autograd import Upsample2xWithBoundary3d as Interp2xBoundary3dFunction  # noqa: F401
from ._autograd import Upsample2xWithBoundary3dModule as Interp2xBoundary3d  # noqa: F401
