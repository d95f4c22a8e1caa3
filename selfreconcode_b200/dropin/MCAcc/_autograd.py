"""Autograd plumbing of the two differentiable grid ops of the path, written once for both:

* trilinear sampling with border padding (align_corners=False) -- differentiable TWICE, because the
  training losses back-propagate through the Jacobian of the skinning field
  (reference: MCAcc/grid_sampler_mine.py:8-65, model/Deformer.py:208-211);
* the (2n-1) upsample + boundary flag of Seg3dLossless (reference: MCAcc/interp2x_boundary3d.py:8-29).

The kernels are reached through the extension-module shims (`GridSamplerMine`, `interp2x_boundary3d`,
resolved at call time so that tests can substitute a backend); this file only wires them into
torch.autograd.  `mode` / `padding_mode` are accepted for signature compatibility with
F.grid_sample; anything but bilinear + border + align_corners=False is refused.
"""
import importlib

import torch.nn as nn
from torch.autograd import Function

_BILINEAR, _BORDER = 0, 1


def _sampler():
    return importlib.import_module("GridSamplerMine")


def _upsampler():
    return importlib.import_module("interp2x_boundary3d")


class _SampleVjp(Function):
    """(volume, grid, cotangent) -> (d volume, d grid): the first-order backward as an op of its own, so
    that autograd can differentiate it again (the kernel `dbackward` is its VJP)."""

    @staticmethod
    def forward(ctx, volume, grid, cotangent):
        ctx.save_for_backward(volume, grid, cotangent)
        g_volume, g_grid = _sampler().backward(volume, grid, cotangent, _BILINEAR, _BORDER)
        return g_volume, g_grid

    @staticmethod
    def backward(ctx, gg_volume, gg_grid):
        volume, grid, cotangent = ctx.saved_tensors
        d_volume, d_grid, d_cotangent = _sampler().dbackward(gg_volume.contiguous(), gg_grid.contiguous(), volume,
                                                             grid, cotangent, _BILINEAR, _BORDER)
        return d_volume, d_grid, d_cotangent


class TrilinearBorderSample3d(Function):
    """out[n,c,p] = trilinear(volume[n,c], grid[n,p]) with border padding, align_corners=False."""

    @staticmethod
    def forward(ctx, volume, grid, mode='bilinear', padding_mode='border', align_corners=False):
        if align_corners:
            raise NotImplementedError("align_corners=True is not implemented (nor used by the reference)")
        if mode != 'bilinear' or padding_mode != 'border':
            raise NotImplementedError("only mode='bilinear', padding_mode='border' is implemented")
        ctx.save_for_backward(volume, grid)
        return _sampler().forward(volume, grid, _BILINEAR, _BORDER)

    @staticmethod
    def backward(ctx, cotangent):
        volume, grid = ctx.saved_tensors
        g_volume, g_grid = _SampleVjp.apply(volume, grid, cotangent)
        return g_volume, g_grid, None, None, None


class Upsample2xWithBoundary3d(Function):
    """[B,C,d,h,w] -> ([B,C,2d-1,2h-1,2w-1] values, bool boundary mask); only the values carry gradient."""

    @staticmethod
    def forward(ctx, coarse, balance_value):
        fine, boundary = _upsampler().forward(coarse.contiguous(), balance_value)
        ctx.mark_non_differentiable(boundary)
        return fine, boundary

    @staticmethod
    def backward(ctx, g_fine, _g_boundary):
        return _upsampler().backward(g_fine.contiguous()), None


class Upsample2xWithBoundary3dModule(nn.Module):
    def __init__(self, balance_value=0.5):
        super().__init__()
        self.balance_value = balance_value

    def forward(self, input):
        return Upsample2xWithBoundary3d.apply(input, self.balance_value)
