"""Drop-in for the reference's pybind module `interp2x_boundary3d`
(MCAcc/cuda/interp2x_boundary3d.cpp:17-36)."""
import torch

from selfreconcode_b200 import ops as _ops


def _check(x, name):
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not x.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if x.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (the reference also instantiates float64; the "
                           "Seg3dLossless path only uses float32)" % name)


def forward(input, balance_value):
    """forward (CUDA): [B,C,d,h,w] -> [output [B,C,2d-1,2h-1,2w-1], is_boundary bool]."""
    _check(input, "input")
    out, bnd = _ops.interp2x3d_forward(input.detach(), balance_value)
    return [out, bnd]


def backward(grad_output):
    """backward (CUDA): adjoint of forward wrt input."""
    _check(grad_output, "grad_output")
    return _ops.interp2x3d_backward(grad_output.detach())
