"""OptimNetwork.propagateTmpPsGrad against tests/golden/propagate.npz (made by the reference's own
method, oracle/make_golden.py).  CPU: host logic with the oracle's 3x3 inverse standing in for the
kernel (checker only); GPU: the shipped path (fused grad f / Jacobian kernels + sr_minv3x3)."""
import types

import numpy as np
import pytest
import torch

import helpers as H


def grad_digest(grad, idx):
    flat = grad.detach().double().cpu().reshape(-1)
    r = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(9000 + idx), dtype=torch.float64)
    stride = max(1, flat.numel() // 128)
    return np.concatenate([[flat.norm().item(), (flat * r).sum().item()], flat[::stride][:128].numpy()])


class OracleSdf(torch.nn.Module):
    """CPU stand-in for the drop-in ImplicitNetwork (which has no CPU path): same parameters,
    evaluated by the oracle's restatement -- lets the host logic of propagateTmpPsGrad run here."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x, ratio):
        from oracle import oracle as O
        net = self.net
        params = [(getattr(net, "lin%d" % l).weight_v, getattr(net, "lin%d" % l).weight_g,
                   getattr(net, "lin%d" % l).bias) for l in range(net.num_layers - 1)]
        return O.sdf_forward(params, x, 6, ratio["sdfRatio"], skip_in=(4,), d_out=1)[0]


class OracleDeformer(torch.nn.Module):
    def __init__(self, comp, gd):
        super().__init__()
        self.comp = comp
        self.gd = gd

    def forward(self, ps, defconds, batch_inds, ratio=None):
        from oracle import oracle as O
        tr, sk = self.comp.defs
        tparams = [(getattr(tr, "lin%d" % l).weight, getattr(tr, "lin%d" % l).bias)
                   for l in range(tr.num_layers - 1)]
        poses, trans = defconds[1]
        A, _ = O.bone_transforms(poses, sk.Js, H.SMPL_PARENTS, sk.init_pose)
        lbs = dict(ws=sk.ws, bmin=sk.b_min.view(-1), bmax=sk.b_max.view(-1), A=A, trans=trans)
        return O.composite_deform(tparams, 6, ratio["deformerRatio"], defconds[0], lbs, ps, batch_inds)[0]


def run_case(case, device):
    H.dropin()
    from model.Deformer import CompositeDeformer
    from model.optim import OptimNetwork
    from model.CameraMine import RectifiedPerspectiveCameras
    g = H.golden("propagate.npz")
    gd = H.golden("deform.npz")
    sdf = H.build_sdf_full(H.golden("sdf_full.npz")).to(device)
    comp = CompositeDeformer([H.build_translator(gd), H.build_skinner(gd)]).to(device)
    sdf_mod, comp_mod = sdf, comp
    if device == "cpu":
        sdf_mod, comp_mod = OracleSdf(sdf), OracleDeformer(comp, gd)
    opt = case == "optcam"
    t = lambda k, src=g: torch.from_numpy(src[k]).to(device)
    cam_t = [t(k).clone().requires_grad_(opt) for k in ("focals", "pps", "Rs", "Ts")]
    cond_t = [t(k, gd).clone().requires_grad_(True) for k in ("poses", "trans", "dcond")]
    Hh, Ww = int(g["H"]), int(g["W"])

    class FakeData:
        def get_grad_parameters(self, fids, dev):
            return cond_t[0], cond_t[1], cond_t[2], None

        def get_camera_parameters(self, n, dev):
            return cam_t[0], cam_t[1], cam_t[2], cam_t[3], Hh, Ww

    cam0 = RectifiedPerspectiveCameras(*[c.detach().cpu() for c in cam_t], image_size=[(Ww, Hh)])
    assert abs(cam0.angThreshold(0.5) - float(g["angthr"])) < 1e-7
    holder = types.SimpleNamespace(rasterizer=types.SimpleNamespace(cameras=cam0))
    on = OptimNetwork(sdf_mod, comp_mod, None, holder, None, conf=None)
    on.dataset = FakeData()
    on.TmpPs = t("tmpps").clone().requires_grad_(True)
    on.TmpPs.grad = t("grad_l_p").clone()
    col, row = t("col"), t("row")
    pix = torch.cat([col.view(-1, 1), row.view(-1, 1), torch.ones_like(col.view(-1, 1))], dim=-1).float()
    camg = RectifiedPerspectiveCameras(*cam_t, image_size=[(Ww, Hh)])
    on.rays = camg.view_rays(pix)
    np.testing.assert_allclose(on.rays.detach().cpu().numpy(), g[case + "_rays"], atol=2e-6)
    on.col_inds, on.row_inds, on.batch_inds = col, row, t("batch_inds")
    on.propagateTmpPsGrad(torch.arange(cond_t[0].shape[0], device=device), H.RATIO)
    assert tuple(on.info["invInfo"]) == tuple(g[case + "_invinfo"])
    named = [("sdf." + k, q) for k, q in sorted(sdf.named_parameters())] + \
            [("def." + k, q) for k, q in sorted(comp.named_parameters())] + \
            list(zip(("poses", "trans", "dcond"), cond_t))
    if opt:
        named += list(zip(("focals", "pps", "Rs", "Ts"), cam_t))
    seen = 0
    for i, (k, q) in enumerate(named):
        key = case + "__" + k
        if key not in g.files:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0, k
            continue
        assert q.grad is not None, k
        got, want = grad_digest(q.grad, i), g[key]
        scale = max(np.abs(want[2:]).max(), want[0] / np.sqrt(max(q.numel(), 1)), 1e-12)
        assert abs(got[0] - want[0]) <= 2e-4 * want[0] + 1e-9, (k, got[0], want[0])
        assert np.abs(got[2:] - want[2:]).max() <= 2e-4 * scale + 1e-9, (k, np.abs(got[2:] - want[2:]).max(), scale)
        seen += 1
    assert seen >= 20 + (4 if opt else 0)


def test_camera_matches_the_reference_class():
    """RectifiedPerspectiveCameras.project / cam_pos / view_rays / angThreshold (model/CameraMine.py:129-170)
    against values produced by the reference's own methods (oracle/make_golden.py)."""
    H.dropin()
    from model.CameraMine import RectifiedPerspectiveCameras
    g = H.golden("propagate.npz")
    t = lambda k: torch.from_numpy(g[k])
    cam = RectifiedPerspectiveCameras(t("focals"), t("pps"), t("Rs"), t("Ts"), image_size=[(int(g["W"]), int(g["H"]))])
    np.testing.assert_allclose(cam.project(t("cam_project_in")).numpy(), g["cam_project"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(cam.cam_pos().numpy(), g["cam_pos"], rtol=0, atol=1e-7)
    assert abs(cam.angThreshold(0.5) - float(g["angthr"])) < 1e-7
    pix = torch.stack([t("col"), t("row"), torch.ones_like(t("col"))], dim=1).float()
    np.testing.assert_allclose(cam.view_rays(pix).numpy(), g["fixedcam_rays"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("case", ["fixedcam", "optcam"])
def test_propagate_host_logic_cpu(case, monkeypatch):
    H.dropin()
    import model.optim as optim
    from oracle import oracle as O
    monkeypatch.setattr(optim, "Fast3x3Minv", lambda ms: O.minv3x3(ms))
    run_case(case, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["fixedcam", "optcam"])
def test_propagate_gpu(case):
    run_case(case, "cuda")
