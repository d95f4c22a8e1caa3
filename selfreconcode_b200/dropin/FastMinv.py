"""Drop-in for the reference's pybind module `FastMinv` (FastMinv/M3x3Inv.cpp:12-63).

Same names, argument meaning and error behaviour: non-CUDA / non-contiguous / wrong dtype
inputs raise RuntimeError (the reference's AT_ASSERTM, M3x3Inv.cpp:4-6,14-15); outputs are
fresh tensors with requires_grad=False; `[invs, checks]` is returned as a list.
Runs on the caller's current CUDA stream (the reference used the legacy default stream).
"""
import torch

from selfreconcode_b200 import ops as _ops


def _check_input(x, name):
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not x.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def Fast3x3Minv(ms):
    """fast batch 3x3 matrix inversion (CUDA): ms [N,3,3] -> [invs [N,3,3], checks [N] bool]."""
    _check_input(ms, "ms")
    if ms.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rs must be a float/double tensor")
    invs, checks = _ops.minv3x3(ms.detach())
    return [invs, checks]


def Fast3x3Minv_backward(grads, invs):
    """fast batch 3x3 matrix inversion backward (CUDA): -(C^T G C^T), C = invs."""
    _check_input(grads, "grads")
    _check_input(invs, "invs")
    if grads.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("grads must be a float/double tensor")
    if invs.dtype != grads.dtype:
        raise RuntimeError("invs must have same type with grads")
    return _ops.minv3x3_backward(grads.detach(), invs.detach())
