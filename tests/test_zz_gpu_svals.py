"""Singular values of 3x3 Jacobians on the device (csrc/svals3x3.cu) against the values the
unmodified reference's `torch.svd` produced for the same matrices (tests/golden/boundary.npz: svd_J / svd_S, written by
oracle/make_golden_r2.py; model/network.py:573-575) and against float64 `torch.linalg.svdvals` with its autograd."""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


def test_singular_values_match_reference_svd(cuda_dev):
    from selfreconcode_b200 import ops
    t = golden("boundary.npz")
    J = torch.from_numpy(t["svd_J"]).float()
    S, V = ops.svals3x3(J.to(cuda_dev), want_v=True)
    S = S.cpu()
    ref = torch.from_numpy(t["svd_S"]).float()
    s64 = torch.linalg.svdvals(J.double())
    print("svals3x3: max |S - torch.svd fixture| %.2e, max |S - float64| %.2e"
          % ((S - ref).abs().max().item(), (S.double() - s64).abs().max().item()))
    assert torch.all(S[:, 0] >= S[:, 1]) and torch.all(S[:, 1] >= S[:, 2])           # torch.svd's order
    assert torch.allclose(S, ref, rtol=1e-5, atol=5e-6)
    assert torch.allclose(S.double(), s64, rtol=1e-5, atol=1e-6)
    # V: right singular vectors (columns), orthonormal, J^T J v_i = s_i^2 v_i
    Vd = V.cpu().double()
    eye = torch.eye(3, dtype=torch.float64).expand_as(Vd)
    assert (Vd.transpose(1, 2) @ Vd - eye).abs().max().item() < 1e-5
    A = J.double().transpose(1, 2) @ J.double()
    assert ((A @ Vd) - Vd * (S.double() ** 2).unsqueeze(1)).abs().max().item() < 1e-5


def test_singular_values_backward_matches_autograd(cuda_dev):
    from selfreconcode_b200 import ops
    t = golden("boundary.npz")
    J = torch.from_numpy(t["svd_J"]).float()
    g = torch.Generator().manual_seed(4)
    gS = torch.randn(J.shape[0], 3, generator=g)
    Jd = J.to(cuda_dev)
    S, V = ops.svals3x3(Jd, want_v=True)
    gJ = ops.svals3x3_backward(Jd, S, V, gS.to(cuda_dev)).cpu()
    J64 = J.double().requires_grad_(True)
    (torch.linalg.svdvals(J64) * gS.double()).sum().backward()
    err = (gJ.double() - J64.grad).abs().max().item()
    print("svals3x3 backward: max abs err %.2e (max |g| %.2e)" % (err, J64.grad.abs().max().item()))
    assert err < 1e-4 * J64.grad.abs().max().item() + 1e-6
