"""Reference names (MCAcc/grid_sampler_mine.py:8,46) for the twice-differentiable sampler in _autograd.py."""
from ._autograd import TrilinearBorderSample3d as GridSamplerMine3dFunction  # noqa: F401
from ._autograd import _SampleVjp as GridSamplerMine3dBackwardFunction  # noqa: F401
