"""Autograd wrappers giving the 3-D grid sampler a double backward
(reference: MCAcc/grid_sampler_mine.py:8-65), over the drop-in GridSamplerMine op."""
import torch
from torch.autograd import Function

import GridSamplerMine


class GridSamplerMine3dFunction(Function):
    @staticmethod
    def forward(ctx, input, grid, mode='bilinear', padding_mode='border', align_corners=False):
        ctx.save_for_backward(input, grid)
        if align_corners == True:
            raise NotImplementedError
        return GridSamplerMine.forward(input, grid, 0, 1)

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        o0, o1 = GridSamplerMine3dBackwardFunction.apply(input, grid, grad_output)
        return o0, o1, None, None, None


class GridSamplerMine3dBackwardFunction(Function):
    @staticmethod
    def forward(ctx, input, grid, grad_output):
        ctx.save_for_backward(input, grid, grad_output)
        return GridSamplerMine.backward(input, grid, grad_output, 0, 1)

    @staticmethod
    def backward(ctx, grad_output_input, grad_output_grid):
        input, grid, grad_output = ctx.saved_tensors
        o0, o1, o2 = GridSamplerMine.dbackward(grad_output_input.contiguous(),
                                               grad_output_grid.contiguous(), input, grid,
                                               grad_output, 0, 1)
        return o0, o1, o2
