"""Diagnostic: which loss term of the optimisation step carries the SDF-gradient difference between the tensor-core
training engine and the torch-autograd twin (same GPU, same modules, same inputs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import test_optim_step_gpu as T  # noqa: E402


def run(net, data, rays, fids, fused, conf, propagate):
    from selfreconcode_b200 import train_ops
    dev = "cuda"
    N, Hh, Ww = fids.numel(), data.H, data.W
    train_ops.TC_TRAIN_ENABLED = fused
    net.conf = conf
    g = torch.Generator().manual_seed(5)
    img = (torch.rand(N, Hh, Ww, 3, generator=g) * 2 - 1).to(dev)
    nrm = torch.nn.functional.normalize(torch.randn(N, Hh, Ww, 3, generator=g), dim=-1).to(dev)
    mods = {"sdf": net.sdf, "def": net.deformer, "rn": net.netRender, "data": data}
    for m in mods.values():
        for q in m.parameters():
            q.grad = None
    torch.manual_seed(9)
    bi, ri, ci = rays["batch_inds"].to(dev), rays["rows"].to(dev), rays["cols"].to(dev)
    loss = net.forward_rays({"img": img, "normal": nrm}, bi, ri, ci, rays["init_pts"].to(dev).clone(), H.RATIO, fids)
    loss.backward()
    gp = net.TmpPs.grad.detach().clone() if net.TmpPs.grad is not None else torch.zeros_like(net.TmpPs)
    if propagate:
        for m in mods.values():
            for q in m.parameters():
                q.grad = None
        net.propagateTmpPsGrad(fids, H.RATIO)
    out = {}
    for name, m in mods.items():
        for k, q in m.named_parameters():
            if q.grad is not None and float(q.grad.abs().max()) > 0:
                out[name + "." + k] = q.grad.detach().clone()
    train_ops.TC_TRAIN_ENABLED = True
    return loss.item(), gp, out


def main():
    from selfreconcode_b200 import synth
    net, data, rays, fids = T.build()
    base = dict(grad_weight=0.0, color_weight=0.0, normal_weight=0.0)
    cases = {
        "eikonal only": (dict(base, grad_weight=0.1), False),
        "colour only": (dict(base, color_weight=0.5), False),
        "normal only (unweighted)": (dict(base, normal_weight=0.1, weighted_normal=False), False),
        "normal only (weighted)": (dict(base, normal_weight=0.1, weighted_normal=True), False),
        "def_regu only": (dict(base, def_regu=dict(weight=2.0, c=0.5)), False),
        "offset only": (dict(base, offset_weight=0.05), False),
        "propagate only (after colour)": (dict(base, color_weight=0.5), True),
    }
    for name, (cf, prop) in cases.items():
        conf = synth.Conf(**cf)
        la, ga, oa = run(net, data, rays, fids, False, conf, prop)
        lf, gf, of = run(net, data, rays, fids, True, conf, prop)
        groups = {}
        for k in oa:
            e = H.norm_err(of[k].cpu().numpy(), oa[k].cpu().numpy()) if k in of else float("nan")
            grp = k.split(".")[0]
            if e > groups.get(grp, (0, ""))[0]:
                groups[grp] = (e, k)
        missing = [k for k in oa if k not in of] + ["+" + k for k in of if k not in oa]
        print("%-32s loss %.6f / %.6f  dL/dp %.1e  %s  %s" % (
            name, la, lf, H.norm_err(gf.cpu().numpy(), ga.cpu().numpy()) if float(ga.abs().max()) > 0 else 0.0,
            {g: "%.1e (%s)" % v for g, v in groups.items()}, ("MISSING " + str(missing[:4])) if missing else ""))


if __name__ == "__main__":
    main()
