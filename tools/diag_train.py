"""Diagnostic: SDF parameter gradients of single loss terms, tensor-core training engine vs torch autograd (fp64 on
GPU), per layer.  Narrows down which output / row type carries a gradient error."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_sdf_full, golden, norm_err  # noqa: E402


def ref_forward4(sdf, p):
    """forward-mode SDF in fp64 torch ops (value + 3 tangents), differentiable."""
    from selfreconcode_b200 import train_ops as T
    P = p.shape[0]
    x0 = T.embed_rows(p, 6, [1.0] * 6, 4, ld=39)
    L = sdf.num_layers - 1
    x = x0
    for l in range(L):
        lin = getattr(sdf, "lin%d" % l)
        W = lin.weight_v * (lin.weight_g.view(-1, 1) / lin.weight_v.norm(dim=1, keepdim=True))
        if l in sdf.skip_in:
            x = torch.cat([x, x0], 1) / np.sqrt(2)
        z = (x @ W.t()).view(P, 4, -1)
        zv = z[:, 0] + lin.bias
        if l < L - 1:
            a = torch.nn.functional.softplus(zv, beta=100)
            d = torch.sigmoid(100 * zv)
            x = torch.cat([a.unsqueeze(1), d.unsqueeze(1) * z[:, 1:]], 1).reshape(P * 4, -1)
        else:
            x = torch.cat([zv.unsqueeze(1), z[:, 1:]], 1)
    return x[:, 0, :1], x[:, 1:, 0], x[:, 0, 1:]


def main():
    dev = torch.device("cuda:0")
    g = golden("train_step.npz")
    torch.manual_seed(0)
    base = torch.from_numpy(g["tmpps"])
    pts = torch.cat([base, base + 0.01 * torch.randn_like(base), torch.rand(600, 3) * 3.6 - 1.8]).to(dev)
    sdf32 = build_sdf_full(golden("sdf_full.npz")).to(dev)
    sdf64 = build_sdf_full(golden("sdf_full.npz")).to(dev).double()
    R = torch.randn(pts.shape[0], 256, device=dev)
    R3 = torch.randn(pts.shape[0], 3, device=dev)
    losses = {
        "f (value rows, cot on f)": lambda f, gf, ft: f.sum(),
        "feat (value rows, cot on the 256 features)": lambda f, gf, ft: (ft * R.to(ft.dtype)).sum() if ft is not None else None,
        "grad f linear (tangent rows, cot R3)": lambda f, gf, ft: (gf * R3.to(gf.dtype)).sum(),
        "eikonal ((|grad f|-1)^2)": lambda f, gf, ft: ((gf.norm(2, dim=-1) - 1) ** 2).mean(),
    }
    for want_feat in (True, False):
        for name, fn in losses.items():
            if not want_feat and name.startswith("feat"):
                continue
            sdf32.zero_grad()
            sdf64.zero_grad()
            p32 = pts.clone().requires_grad_(True)
            out = sdf32.forward_train(p32, 1.0, want_grad=True, want_feat=want_feat)
            fn(*out).backward()
            p64 = pts.double().clone().requires_grad_(True)
            o64 = ref_forward4(sdf64, p64)
            fn(o64[0], o64[1], o64[2] if want_feat else None).backward()
            errs = []
            for (k, a), (_, b) in zip(sdf32.named_parameters(), sdf64.named_parameters()):
                if b.grad is None or float(b.grad.abs().max()) == 0:
                    continue
                errs.append((k, norm_err(a.grad.cpu().numpy(), b.grad.cpu().numpy())))
            ep = norm_err(p32.grad.cpu().numpy(), p64.grad.cpu().numpy())
            worst = max(errs, key=lambda kv: kv[1])
            print("want_feat=%s  %-48s dL/dp %.1e  worst param %s %.1e  per layer v: %s"
                  % (want_feat, name, ep, worst[0], worst[1],
                     " ".join("%.0e" % e for k, e in errs if k.endswith("weight_v"))))
    # value-only sweeps (1 row per point), as the template-vertex term |f(TmpVs)| and the implicit-differentiation VJP use
    def ref_f(sdf, p):
        from selfreconcode_b200 import train_ops as T
        x0 = T.embed_rows(p, 6, [1.0] * 6, 1, ld=39)
        x = x0
        L = sdf.num_layers - 1
        for l in range(L):
            lin = getattr(sdf, "lin%d" % l)
            W = lin.weight_v * (lin.weight_g.view(-1, 1) / lin.weight_v.norm(dim=1, keepdim=True))
            if l in sdf.skip_in:
                x = torch.cat([x, x0], 1) / np.sqrt(2)
            x = x @ W.t() + lin.bias
            if l < L - 1:
                x = torch.nn.functional.softplus(x, beta=100)
        return x[:, :1]

    for npts in (pts.shape[0], 2472, 174, 3614):
        pp = torch.cat([pts] * 4)[:npts]
        cot = torch.randn(npts, 1, device=dev)
        for name, fn in (("mean |f|", lambda f: f.abs().mean()), ("VJP random cot", lambda f: (f * cot.to(f.dtype)).sum())):
            sdf32.zero_grad()
            sdf64.zero_grad()
            p32 = pp.clone().requires_grad_(True)
            fn(sdf32.forward_train(p32, 1.0, want_grad=False, want_feat=False)[0]).backward()
            p64 = pp.double().clone().requires_grad_(True)
            fn(ref_f(sdf64, p64)).backward()
            errs = [(k, norm_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()))
                    for (k, a), (_, b) in zip(sdf32.named_parameters(), sdf64.named_parameters())
                    if b.grad is not None and float(b.grad.abs().max()) > 0]
            worst = max(errs, key=lambda kv: kv[1])
            print("1 row/pt, %5d pts  %-16s dL/dp %.1e  worst %s %.1e  v: %s" % (
                npts, name, norm_err(p32.grad.cpu().numpy(), p64.grad.cpu().numpy()), worst[0], worst[1],
                " ".join("%.0e" % e for k, e in errs if k.endswith("weight_v"))))
    # 4 rows/pt with an odd number of row tiles (3614 points -> 113 tiles)
    pp = torch.cat([pts] * 4)[:3614]
    sdf32.zero_grad()
    sdf64.zero_grad()
    p32 = pp.clone().requires_grad_(True)
    eik = lambda gf: ((gf.norm(2, dim=-1) - 1) ** 2).mean()
    eik(sdf32.forward_train(p32, 1.0, True, False)[1]).backward()
    p64 = pp.double().clone().requires_grad_(True)
    eik(ref_forward4(sdf64, p64)[1]).backward()
    errs = [(k, norm_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()))
            for (k, a), (_, b) in zip(sdf32.named_parameters(), sdf64.named_parameters())
            if b.grad is not None and float(b.grad.abs().max()) > 0]
    print("eikonal, 3614 pts (113 row tiles): dL/dp %.1e  v: %s" % (
        norm_err(p32.grad.cpu().numpy(), p64.grad.cpu().numpy()), " ".join("%.0e" % e for k, e in errs if k.endswith("weight_v"))))
    # forward accuracy by row type
    with torch.no_grad():
        f, gf, ft = sdf32.forward_train(pts, 1.0, True, True)
        f64, gf64, ft64 = ref_forward4(sdf64, pts.double())
    print("forward: f %.1e  grad f %.1e  feat %.1e" % (norm_err(f.cpu().numpy(), f64.cpu().numpy()),
                                                      norm_err(gf.cpu().numpy(), gf64.cpu().numpy()),
                                                      norm_err(ft.cpu().numpy(), ft64.cpu().numpy())))


if __name__ == "__main__":
    main()
