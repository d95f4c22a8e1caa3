"""Host-side logic of the tensor-core engine that needs no GPU: how a chain of layer launches is split between the
whole-sweep entry point (sr_tc_sweep) and per-layer calls (sr_tc_linear), and the struct layouts the C ABI reads."""
import ctypes as C

import torch

from selfreconcode_b200 import _lib, ops

NONE, SOFTPLUS, RELU, TANH = 0, 1, 2, 3


class FakeLib:
    def __init__(self):
        self.calls = []

    def sr_tc_sweep(self, arr, n, M, ch, m_dev, stream):
        self.calls.append(("sweep", n, [(arr[i].act, arr[i].mul_act, bool(arr[i].mul_tiles), arr[i].N, arr[i].K)
                                        for i in range(n)]))
        return 0

    def sr_tc_linear(self, *a):
        self.calls.append(("linear", a[4], a[5], a[7]))       # N, K, act
        return 0


def _chain(acts, mul=None, width=512, dstash_at=None):
    t = torch.zeros(8)
    steps = []
    for i, a in enumerate(acts):
        last = i == len(acts) - 1
        kw = dict(A=t, W=t, bias=t, N=width if not last else 1, K=width, n_valid=width if not last else 1, act=a,
                  A_next=None if last else t, K_next=0 if last else width, out=t if last else None)
        if mul is not None and mul[i] is not None:
            kw.update(mul_tiles=t, mul_K=width, mul_act=mul[i], act=NONE)
        if dstash_at == i:
            kw.update(dstash=t)
        steps.append(ops._tc_step(**kw))
    return steps


def _run(monkeypatch, steps, sweep=True):
    lib = FakeLib()
    monkeypatch.setattr(ops, "_stream", lambda: C.c_void_p(0))
    monkeypatch.setattr(ops, "TC_SWEEP", sweep)
    n = ops._run_steps(lib, steps, 4096, 1, None)
    return n, lib.calls


def test_sweep_grouping(monkeypatch):
    # SDF forward: eight softplus layers and the plain output layer -> one sweep of nine steps
    n, calls = _run(monkeypatch, _chain([SOFTPLUS] * 8 + [NONE]))
    assert n == 9 and [c[0] for c in calls] == ["sweep"] and calls[0][1] == 9
    # renderer: ReLU layers + tanh head -> the ReLU run as a sweep, the head as its own launch
    n, calls = _run(monkeypatch, _chain([RELU] * 4 + [TANH]))
    assert n == 4 and [c[0] for c in calls] == ["sweep", "linear"] and calls[1][3] == TANH
    # reverse sweep: act' multiplies on every step but the last (input gradient)
    n, calls = _run(monkeypatch, _chain([NONE] * 9, mul=[SOFTPLUS] * 8 + [None]))
    assert n == 9 and all(m for (_, _, m, _, _) in calls[0][2][:8]) and not calls[0][2][8][2]
    # a training forward (fp32 act' stash requested) never takes the sweep entry point
    n, calls = _run(monkeypatch, _chain([SOFTPLUS] * 3 + [NONE], dstash_at=0))
    assert n == 0 and [c[0] for c in calls] == ["linear"] * 4
    # unsupported activation first, or a single step: per-layer launches
    assert _run(monkeypatch, _chain([TANH, TANH, NONE]))[0] == 0
    assert _run(monkeypatch, _chain([NONE]))[0] == 0
    # more than twelve steps do not fit the kernel's parameter block
    assert _run(monkeypatch, _chain([RELU] * 13 + [NONE]))[0] == 0
    # switched off: identical launches, one per layer
    n, calls = _run(monkeypatch, _chain([SOFTPLUS] * 8 + [NONE]), sweep=False)
    assert n == 0 and len(calls) == 9


def test_struct_layouts_match_header():
    """ctypes mirrors of the structs in include/selfrecon_b200.h: sizes follow from the field lists there
    (8 pointers + 12 ints + 2 floats; 7 pointers + 4 ints; 4 pointers + 4 ints)."""
    assert C.sizeof(_lib.TcStep) == 8 * 8 + 12 * 4 + 2 * 4
    assert C.sizeof(_lib.WnLayer) == 7 * 8 + 4 * 4
    assert C.sizeof(_lib.TcLayer) == 4 * 8 + 4 * 4
    assert _lib.TcStep.N.offset == 64 and _lib.TcStep.scale.offset == 112
    assert _lib.WnLayer.n.offset == 56


def test_effective_weights_cache_and_fallback():
    """ImplicitNetwork._effective / shared_weights (one weight-norm sub-graph per step) and train_ops.weight_norm_all
    off the GPU: the element-wise torch form, same tensors inside one shared context, no state left behind."""
    from helpers import dropin
    dropin()
    from selfreconcode_b200 import synth, train_ops as T
    sdf = synth.make_sdf()
    L = sdf.num_layers - 1
    lins = [getattr(sdf, "lin%d" % l) for l in range(L)]
    ref = [T.weight_norm_eff(lin.weight_v, lin.weight_g) for lin in lins]
    assert all(torch.equal(a, b) for a, b in zip(T.weight_norm_all(lins), ref))
    assert torch.equal(sdf._effective(3), ref[3])                 # no context: computed on demand
    with sdf.shared_weights():
        got = [sdf._effective(l) for l in range(L)]
        assert all(torch.equal(a, b) for a, b in zip(got, ref))
        assert sdf._effective(2) is got[2]                        # one sub-graph: the same tensor object
    assert sdf._weff is None
    # gradients flow to both parameters through the shared sub-graph
    with sdf.shared_weights():
        (sdf._effective(1).sum() + sdf._effective(1).pow(2).sum()).backward()
    assert lins[1].weight_v.grad is not None and lins[1].weight_g.grad is not None


def test_weight_norm_backward_formula():
    """The closed form csrc/weight_norm.cu implements (gg = <gw, v>/||v||, gv = g/||v|| (gw - v <gw, v>/||v||^2))
    against torch autograd of g v/||v|| in float64."""
    g0 = torch.Generator().manual_seed(2)
    v = torch.randn(7, 13, generator=g0, dtype=torch.float64, requires_grad=True)
    g = torch.randn(7, 1, generator=g0, dtype=torch.float64, requires_grad=True)
    gw = torch.randn(7, 13, generator=g0, dtype=torch.float64)
    (g * v / v.norm(dim=1, keepdim=True) * gw).sum().backward()
    nrm = v.detach().norm(dim=1, keepdim=True)
    s = (gw * v.detach()).sum(1, keepdim=True)
    gg = s / nrm
    gv = g.detach() / nrm * (gw - v.detach() * s / nrm ** 2)
    assert torch.allclose(gg, g.grad, rtol=1e-12, atol=1e-14)
    assert torch.allclose(gv, v.grad, rtol=1e-12, atol=1e-14)


def test_bench_parity_report_counts():
    """bench.parity_report on hand-made inputs: the bar is applied on decision-insensitive rays only, sign differences
    are split by the oracle's own |f| band, the mesh comparison is exact."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    P = 6
    pts = torch.linspace(0.1, 1.0, P * 3).view(P, 3)
    rgb = torch.linspace(-0.5, 0.5, P * 3).view(P, 3)
    conv = torch.tensor([1, 1, 0, 1, 0, 1], dtype=torch.bool)
    cpu = {"pts": pts.clone(), "rgb": rgb.clone(), "conv": conv.clone(),
           "sensitive": torch.tensor([0, 0, 0, 0, 1, 1], dtype=torch.bool)}
    gpu = {"pts": pts.clone(), "rgb": rgb.clone(), "conv": conv.clone()}
    gpu["pts"][1, 0] += 1e-2          # insensitive ray over tolerance
    gpu["pts"][5, 2] += 1e-2          # sensitive ray over tolerance: reported, not held against the bar
    gpu["conv"][4] = True             # mask flip on a sensitive ray
    gpu["rgb"][0, 1] += 3e-5          # inside 1e-4 |b| + 1e-4 mean|b|
    grid_c = torch.tensor([[-1.0, 2e-6, 0.5], [0.3, -0.2, -4e-6]]).numpy()
    grid_g = torch.tensor([[-1.0, -1e-6, 0.5], [-0.3, -0.2, -4e-6]])      # one flip inside the band, one outside
    calc = torch.ones(2, 3, dtype=torch.bool)
    faces = np.array([[0, 1, 2], [2, 1, 3]], dtype=np.int64)
    verts = np.zeros((4, 3), dtype=np.float32)
    cpu.update(grid=torch.from_numpy(grid_c), calc=calc, faces=faces, verts=verts)
    gpu.update(grid=grid_g, calc=calc.clone(), faces=torch.from_numpy(faces), verts=torch.from_numpy(verts),
               verts_on_oracle_grid=torch.from_numpy(verts), faces_on_oracle_grid=torch.from_numpy(faces[::-1].copy()))
    rep = bench.parity_report(gpu, cpu, band=1e-5)
    assert rep["rays"] == P and rep["rays_decision_sensitive"] == 2
    assert rep["conv_mismatch_all"] == 1 and rep["conv_mismatch_insensitive"] == 0
    assert rep["pts_rays_over_tol_all"] == 2 and rep["pts_rays_over_tol_insensitive"] == 1
    assert rep["rgb_rays_over_tol_insensitive"] == 0
    assert rep["sign_mismatch"] == 2 and rep["sign_mismatch_outside_fp32_band"] == 1
    assert rep["queried_set_mismatch"] == 0 and rep["mc_mesh_identical"] is True
    assert rep["mc_on_oracle_grid_faces_identical"] is False and rep["mc_on_oracle_grid_verts_identical"] is True
