"""Drop-in for the reference's pybind module `GridSamplerMine`
(MCAcc/cuda/GridSamplerMine.cpp:24-103): forward / backward / dbackward of the 3-D trilinear
sampler with border padding and align_corners=False.  Same checks as the reference's
TORCH_CHECKs (:27-69): both tensors defined, same device and dtype, strided layout, 5-D,
grid.size(-1)==3, only interpolation 0 (bilinear) with padding 1 (border).
"""
import torch

from selfreconcode_b200 import ops as _ops


def _check(input, grid, interpolation_mode, padding_mode):
    if input.device != grid.device:
        raise RuntimeError("grid_sampler(): expected input and grid to be on same device, but input "
                           "is on %s and grid is on %s" % (input.device, grid.device))
    if input.dtype != grid.dtype:
        raise RuntimeError("grid_sampler(): expected input and grid to have same dtype, but input "
                           "has %s and grid has %s" % (input.dtype, grid.dtype))
    if input.dim() != 5 or grid.dim() != 5:
        raise RuntimeError("grid_sampler(): expected 5D input and grid with same number of "
                           "dimensions")
    if input.size(0) != grid.size(0):
        raise RuntimeError("grid_sampler(): expected grid and input to have same batch size")
    if grid.size(-1) != 3:
        raise RuntimeError("grid_sampler(): expected grid to have size 3 in last dimension")
    if any(s <= 0 for s in input.shape[2:]):
        raise RuntimeError("grid_sampler(): expected input to have non-empty spatial dimensions")
    if interpolation_mode != 0 or padding_mode != 1:
        raise RuntimeError("GridSamplerMine: only bilinear interpolation (0) with border padding "
                           "(1) is implemented")
    if not input.is_cuda:
        raise RuntimeError("GridSamplerMine: CUDA tensors required")


def forward(input, grid, interpolation_mode=0, padding_mode=1):
    _check(input, grid, interpolation_mode, padding_mode)
    return _ops.grid_sample3d_forward(input.detach(), grid.detach())


def backward(input, grid, grad_output, interpolation_mode=0, padding_mode=1):
    _check(input, grid, interpolation_mode, padding_mode)
    return _ops.grid_sample3d_backward(input.detach(), grid.detach(), grad_output.detach())


def dbackward(grad_output_input, grad_output_grid, input, grid, grad_output, interpolation_mode=0,
              padding_mode=1):
    _check(input, grid, interpolation_mode, padding_mode)
    return _ops.grid_sample3d_dbackward(grad_output_input.detach(), grad_output_grid.detach(),
                                        input.detach(), grid.detach(), grad_output.detach())
