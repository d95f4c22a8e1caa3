"""Math helpers of the hot path (reference: utils/utils.py:8-172, 237-316).

Only what the per-frame optimisation step touches is mirrored here; the mesh / texture /
checkpoint-format helpers of the reference are plain torch and out of scope."""
import numpy as np
import torch
from torch.autograd import Function

from FastMinv import Fast3x3Minv, Fast3x3Minv_backward
from selfreconcode_b200 import ops


class FastDiff3x3MinvFunction(Function):
    @staticmethod
    def forward(ctx, input):
        invs, check = Fast3x3Minv(input.contiguous())
        ctx.save_for_backward(invs, check)
        ctx.mark_non_differentiable(check)
        return invs, check

    @staticmethod
    def backward(ctx, grad_input, grad_check):
        invs, check = ctx.saved_tensors
        return Fast3x3Minv_backward(grad_input.contiguous(), invs), None


def quat2mat(quat):
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2],
                       dim=1).view(quat.size(0), 3, 3)


def annealing_weights(multires, ratio):
    alpha = ratio * multires
    out = []
    for ind in range(multires):
        w = (1. - np.cos(np.pi * min(max(alpha - float(ind), 0.), 1.))) / 2.
        out.extend([w, w])
    return out


def GMRobustError(x, c, square=False):
    if square:
        return 2. * x / (c * c) / (x / (c * c) + 4)
    return 2. * x * x / (c * c) / (x * x / (c * c) + 4)


def smpl_tmp_Apose(init_pose_type=0):
    pose = np.zeros((24, 3))
    leg, arm = {0: (10., 45.), 1: (7., 55.)}[init_pose_type]
    pose[1] = np.array([0, 0, leg / 180. * np.pi])
    pose[2] = np.array([0, 0, -leg / 180. * np.pi])
    pose[16] = np.array([0, 0, -arm / 180. * np.pi])
    pose[17] = np.array([0, 0, arm / 180. * np.pi])
    return pose.astype(np.float32)


def sample_points(pc_input, global_sigma, local_sigma, ratio=6):
    sample_size, dim = pc_input.shape
    sample_local = pc_input + (torch.randn_like(pc_input) * local_sigma)
    if ratio > 0:
        sample_global = (torch.rand(sample_size // ratio, dim, device=pc_input.device)
                         * (global_sigma * 2)) - global_sigma
        return torch.cat([sample_local, sample_global], dim=0)
    return sample_local


def compute_Jacobian(ps, ds, retain_graph, create_graph, allow_unused=False):
    """J[p, i, :] = d ds_i / d ps via three VJPs (autograd path, used when a graph is needed)."""
    rows = []
    ones = torch.ones_like(ds[..., 0])
    for i in range(3):
        keep = True if i < 2 else retain_graph
        out = torch.autograd.grad(ds[..., i], ps, ones, retain_graph=keep,
                                  create_graph=create_graph, allow_unused=allow_unused)
        rows.append(out[0].view(-1, 1, 3))
    return torch.cat(rows, dim=1)


def _fusable(deformer):
    return hasattr(deformer, "_fusable") and deformer._fusable()


def compute_deformed_normals(sdf, deformer, ps, defconds, batch_inds, ratio, phase):
    check = phase in ('train', 'Train')
    if not check and _fusable(deformer) and hasattr(sdf, "forward_fused"):
        # no graph requested: f, grad f, D, dD/dp from the fused forward-mode kernels
        _, onx, _ = sdf.forward_fused(ps, ratio, want_grad=True, want_feat=False)
        ds, grad_d_p, _ = deformer.forward_fused(ps, defconds, batch_inds, ratio, want_jac=True)
    else:
        sdfs = sdf(ps, ratio)
        onx = torch.autograd.grad(sdfs, ps, torch.ones_like(sdfs), retain_graph=check,
                                  create_graph=check)[0]
        ds = deformer(ps, defconds, batch_inds, ratio=ratio)
        grad_d_p = compute_Jacobian(ps, ds, check, check)
    grad_d_p_inv, inv_mask = FastDiff3x3MinvFunction.apply(grad_d_p)
    nx = grad_d_p_inv.transpose(-2, -1).matmul(onx.view(-1, 3, 1)).view(-1, 3)
    n_inv_mask = ~inv_mask
    if n_inv_mask.sum().item() > 0:
        print('unwished error n_inv_mask:(%d:%d)' % (n_inv_mask.sum().item(), n_inv_mask.numel()))
        nnx = torch.zeros_like(nx)
        nnx[inv_mask] = nx[inv_mask]
        nnx[n_inv_mask] = grad_d_p[n_inv_mask].matmul(onx[n_inv_mask].unsqueeze(-1)).view(-1, 3)
        nx = nnx
    nx = nx / nx.norm(dim=1, keepdim=True)
    return nx, ds


def compute_cardinal_rays(deformer, ps, rays, defconds, batch_inds, ratio, phase):
    check = phase in ('train', 'Train')
    if not check and _fusable(deformer):
        ds, grad_d_p, _ = deformer.forward_fused(ps, defconds, batch_inds, ratio, want_jac=True)
    else:
        ds = deformer(ps, defconds, batch_inds, ratio=ratio)
        grad_d_p = compute_Jacobian(ps, ds, check, check)
    grad_d_p_inv, inv_mask = FastDiff3x3MinvFunction.apply(grad_d_p)
    crays = grad_d_p_inv.matmul(rays.view(-1, 3, 1)).view(-1, 3)
    n_inv_mask = ~inv_mask
    if n_inv_mask.sum().item() > 0:
        print('unwished error n_inv_mask:(%d:%d)' % (n_inv_mask.sum().item(), n_inv_mask.numel()))
        ncrays = torch.zeros_like(crays)
        ncrays[inv_mask] = crays[inv_mask]
        ncrays[n_inv_mask] = rays[n_inv_mask].detach()
        crays = ncrays
    crays = crays / crays.norm(dim=1, keepdim=True)
    return crays, ds


def compute_netRender_color(net, ps, ds, ns, vs, features, framefeatures, ratio):
    return net(ps, ns, vs, features, ratio)


def shade_rays(sdf, deformer, netRender, ps, rays, defconds, batch_inds, ratio):
    """Everything the infer loop does with a traced point (model/network.py:356-368): template
    normal grad f / |grad f|, cardinal ray J^-1 v, rendered colour.  -> (normals, crays, rgb), no
    graph.  Large batches with the stock field modules run the SDF / translator sweeps (value +
    3 forward tangents), the pointwise geometry and the rendering network back to back on the
    tensor-core engine (ops.shade_and_render_tc); otherwise the per-op fused kernels are used."""
    P = ps.shape[0]
    stock = (_fusable(deformer) and hasattr(sdf, "fused") and hasattr(netRender, "fused")
             and getattr(netRender, "mode", None) == 'idr' and getattr(netRender, "multires_n", 1) == 0
             and getattr(sdf, "d_out", 0) == 1)
    with torch.no_grad():
        if stock and ops.TC_ENABLED and P >= ops.TC_MIN_POINTS:
            tr, sk = deformer.defs[0], deformer.defs[1]
            full = sdf.fused()
            full.set_pe_weights(sdf._pe_weights(ratio['sdfRatio'] if isinstance(ratio, dict) else ratio))
            poses, trans = defconds[1]
            lbs = sk.lbs_state()
            lbs.set_pose(poses.view(poses.shape[0], 24, 3), trans)
            nfeat = full.desc.layer[full.desc.n_layers - 1].n - 1
            n, cr, rgb, _, _ = ops.shade_and_render_tc(full, tr.fused(ratio), lbs, netRender.fused(ratio), ps, rays,
                                                       batch_inds, defconds[0], nfeat=nfeat)
            return n, cr, rgb
        _, nx, feat = sdf.forward_fused(ps, ratio, want_grad=True, want_feat=True)
        nx = nx / nx.norm(dim=1, keepdim=True)
        crays, defVs = compute_cardinal_rays(deformer, ps, rays, defconds, batch_inds, ratio, 'test')
        rgb = compute_netRender_color(netRender, ps, defVs, nx, crays, feat, None, ratio)
    return nx, crays, rgb
