"""Drop-in for the reference's pybind module `interp2x_boundary2d`
(MCAcc/cuda/interp2x_boundary2d.cpp:17-36); unused by the reference's Python, kept for
import parity."""
import torch

from selfreconcode_b200 import ops as _ops


def _check(x, name):
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not x.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if x.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)


def forward(input, balance_value):
    _check(input, "input")
    out, bnd = _ops.interp2x2d_forward(input.detach(), balance_value)
    return [out, bnd]


def backward(grad_output):
    _check(grad_output, "grad_output")
    return _ops.interp2x2d_backward(grad_output.detach())
