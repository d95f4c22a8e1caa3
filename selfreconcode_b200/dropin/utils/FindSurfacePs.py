"""Surface-point seeding and the ray/surface intersection finder
(reference: utils/FindSurfacePs.py:5-29, 114-163)."""
import torch

from selfreconcode_b200 import ops


def _first_valid(inner):
    """index of the first True along the last (K) axis, K if none (replaces the reference's
    torch_scatter 'min' reduction, FindSurfacePs.py:12-15)."""
    K = inner.shape[-1]
    ar = torch.arange(K, device=inner.device).view(*([1] * (inner.dim() - 1)), K)
    return torch.where(inner, ar, torch.full_like(ar, K)).min(dim=-1)[0]


def FindSurfacePs(TmpVs, TmpFaces, frags):
    N, H, W, K = frags.pix_to_face.shape
    pix_to_face = frags.pix_to_face
    bary = frags.bary_coords
    inner = (bary > 0.0).all(-1) & (pix_to_face >= 0)
    index = _first_valid(inner)
    hit = inner.any(dim=-1)
    batch_inds, row_inds, col_inds = hit.nonzero(as_tuple=True)
    sel = index[hit].view(-1, 1)
    finds = torch.gather(pix_to_face[hit], 1, sel).view(-1)
    finds = finds % TmpFaces.shape[0]  # packed -> per-mesh face index
    ws = torch.gather(bary[hit], 1, sel.view(-1, 1, 1).expand(-1, 1, 3)).view(-1, 3)
    initTmpPs = (TmpVs[TmpFaces[finds].view(-1)].view(-1, 3, 3) * ws[:, :, None]).sum(1)
    return batch_inds, row_inds, col_inds, initTmpPs, finds


def _fused_ok(tmpSdf, deformer):
    return hasattr(tmpSdf, "fused") and hasattr(deformer, "_fusable") and deformer._fusable()


def OptimizeSurfacePs(cam_pos, rays, initTmpPs, batch_inds, tmpSdf, ratio, deformer, defconds,
                      dthreshold=5.e-5, athreshold=0.02, w1=3.05, w2=1., times=5):
    """Damped-Newton search for the point where the camera ray meets the deformed zero level
    set.  Returns (points [P,3] detached, converged [P] bool), like the reference.

    With the fused field modules this is `times+1` back-to-back launches of one persistent
    kernel (csrc/mlp_kernels.cu: trace_kernel) that evaluates f, grad f, D and dD/dp in forward
    mode, tests convergence and updates the unconverged rays, compacting the active list on
    the device -- no autograd graph, no boolean-mask indexing, no host synchronisation."""
    if _fused_ok(tmpSdf, deformer):
        tr, sk = deformer.defs[0], deformer.defs[1]
        if not rays.is_cuda:
            raise RuntimeError("OptimizeSurfacePs: CUDA tensors required (no CPU path)")
        sdf_net = tmpSdf.fused_sdf_only()
        sdf_net.set_pe_weights(tmpSdf._pe_weights(ratio['sdfRatio'] if isinstance(ratio, dict) else ratio))
        dnet = tr.fused(ratio)
        poses, trans = defconds[1]
        lbs = sk.lbs_state()
        lbs.set_pose(poses.view(poses.shape[0], 24, 3), trans)
        pts, conv = ops.trace_surface_points(sdf_net, dnet, lbs, cam_pos, rays, initTmpPs,
                                             batch_inds, defconds[0], dthreshold, athreshold, w1, w2,
                                             times)
        return pts, conv
    return _optimize_generic(cam_pos, rays, initTmpPs, batch_inds, tmpSdf, ratio, deformer, defconds,
                             dthreshold, athreshold, w1, w2, times)


def _optimize_generic(cam_pos, rays, initTmpPs, batch_inds, tmpSdf, ratio, deformer, defconds,
                      dthreshold, athreshold, w1, w2, times):
    """Module-agnostic path for user-supplied field modules (e.g. an identity deformer): the
    reference's algorithm on whatever the modules compute on the GPU."""
    import numpy as np
    if not rays.is_cuda:
        raise RuntimeError("OptimizeSurfacePs: CUDA tensors required (no CPU path)")

    def test(p, r, b):
        with torch.no_grad():
            c1 = tmpSdf(p, ratio).view(-1).abs() < dthreshold
            direct = deformer(p, defconds, b, ratio=ratio) - cam_pos.view(1, 3)
            up = torch.cross(direct, r, dim=1)
            c2 = torch.arcsin(up.norm(dim=1) / direct.norm(dim=1)) * 180. / np.pi < athreshold
            return c1 & c2

    unfinished = ~test(initTmpPs, rays, batch_inds)
    for _ in range(times):
        cur = initTmpPs[unfinished].detach().clone()
        if cur.shape[0] == 0:
            break
        cur.requires_grad_(True)
        b = batch_inds[unfinished]
        loss1 = tmpSdf(cur, ratio).abs().view(-1)
        direct = deformer(cur, defconds, b, ratio=ratio) - cam_pos.view(1, 3)
        up = torch.cross(direct, rays[unfinished], dim=1)
        loss = w1 * loss1 + w2 * (up.norm(dim=1) / direct.norm(dim=1)).abs()
        grad = torch.autograd.grad(loss.sum(), cur)[0]
        cur = (cur + (-loss / (grad * grad).sum(1)).view(-1, 1) * grad).detach()
        initTmpPs[unfinished] = cur
        ok = test(cur, rays[unfinished], b)
        idx = unfinished.nonzero().view(-1)
        unfinished[idx[ok]] = False
    return initTmpPs.detach(), ~unfinished
