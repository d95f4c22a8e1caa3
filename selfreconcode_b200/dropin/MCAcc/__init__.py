"""Drop-in `MCAcc` package: the names the reference's drivers import from it (train.py:117, model/Deformer.py:8)."""
from . import grid_sampler_mine as _sampler_names
from . import seg3d_lossless as _seg3d
from . import utils as _grid_utils

Seg3dLossless = _seg3d.Seg3dLossless
create_grid3D = _grid_utils.create_grid3D
GridSamplerMine3dFunction = _sampler_names.GridSamplerMine3dFunction

__all__ = ["Seg3dLossless", "create_grid3D", "GridSamplerMine3dFunction"]
