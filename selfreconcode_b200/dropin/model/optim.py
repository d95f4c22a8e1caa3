"""Per-step orchestration of the hot path (reference: model/network.py:149-814, OptimNetwork).

What is mirrored, with the reference's names and call conventions:
  * `discretizeSDF`            network.py:292-302  coarse-to-fine SDF grid + marching cubes
  * `infer` (ray part) as `infer_rays`   :342-372  trace + shade every foreground ray of N frames
  * `forward` (ray part) as `forward_rays` :509-644  trace, eikonal, colour and normal losses
  * `propagateTmpPsGrad`       :702-814  implicit differentiation of {f(p)=0, (D(p)-c) x v = 0}
`forward()` / `infer()` themselves start with a pytorch3d mesh / point rasterisation of the deformed
template (network.py:485-505, 317-345) that seeds `FindSurfacePs`; that seed is out of scope
(SURVEY.md section 8f-1), so here they take the seed (batch/row/col indices + start points) from the
caller or from an injected `raster_seed` callable and then run exactly the reference's sequence.
No-grad evaluations run on the fused kernels; losses that need a graph use the modules' autograd
path (same math as torch ops on the GPU).
"""
import numpy as np
import torch
import torch.nn as nn

import utils
from FastMinv import Fast3x3Minv
import MCGpu
from .CameraMine import RectifiedPerspectiveCameras


def _scatter_mean(src, index, n):
    """torch_scatter.scatter(..., reduce='mean', dim_size=n) (network.py:617,637)."""
    s = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add(0, index, src)
    c = torch.zeros(n, dtype=src.dtype, device=src.device).index_add(
        0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
    return s / c.view((-1,) + (1,) * (src.dim() - 1))


def _cross_matrix(v):
    m = torch.zeros(v.shape[0], 3, 3, dtype=v.dtype, device=v.device)
    m[:, 0, 1], m[:, 0, 2] = -v[:, 2], v[:, 1]
    m[:, 1, 0], m[:, 1, 2] = v[:, 2], -v[:, 0]
    m[:, 2, 0], m[:, 2, 1] = -v[:, 1], v[:, 0]
    return m


class OptimNetwork(nn.Module):
    def __init__(self, TmpSdf, Deformer, accEngine, maskRender, netRender, conf=None):
        super().__init__()
        self.conf = conf
        self.sdf = TmpSdf
        self.deformer = Deformer
        self.maskRender = maskRender
        self.netRender = netRender
        self.engine = accEngine
        cams = maskRender.rasterizer.cameras if maskRender is not None else None
        self.angThred = cams.angThreshold(0.5) if cams is not None else 0.02
        self.TmpVs = None
        self.Tmpfs = None
        self.forward_time = 0
        self.remesh_intersect = 30
        self.remesh_time = 0.
        self.next_conf = None
        self.next_train_conf = None
        self.pcRender = None
        self.draw = False
        self.enable_mesh_color = True
        self.sdfShrinkRadius = 0.0
        self.info = {}
        self.TmpPs = None
        self.raster_seed = None   # callable(frame_ids, TmpVs, Tmpfs, defconds, ratio) -> seed dict

    # ---- cameras ----------------------------------------------------------------------------
    def _cameras(self, n, device):
        focals, pps, Rs, Ts, H, W = self.dataset.get_camera_parameters(n, device)
        return RectifiedPerspectiveCameras(focals, pps, Rs, Ts, image_size=[(W, H)]).to(device), H, W

    # ---- network.py:292-302 -------------------------------------------------------------------
    def discretizeSDF(self, ratio, engine=None, balance_value=0.):
        sdf = self.sdf

        def query_func(points):
            with torch.no_grad():
                if hasattr(sdf, "forward_fused"):  # value only: skip the 256-d feature head
                    return sdf.forward_fused(points.reshape(-1, 3), ratio, False, False,
                                             refine_about=balance_value)[0].reshape(1, 1, -1)
                return sdf.forward(points.reshape(-1, 3), ratio).reshape(1, 1, -1)

        if engine is None:
            engine = self.engine
        engine.balance_value = balance_value
        engine.query_func = query_func
        sdfs = engine.forward()
        verts, faces = MCGpu.mc_gpu(sdfs[0, 0].permute(2, 1, 0).contiguous(), engine.spacing_x,
                                    engine.spacing_y, engine.spacing_z, engine.bx, engine.by, engine.bz,
                                    balance_value)
        return verts, faces

    # ---- ray part of infer(): network.py:342-372 ------------------------------------------------
    def infer_rays(self, batch_inds, row_inds, col_inds, initTmpPs, H, W, ratio, frame_ids, chunk=1 << 18):
        device = initTmpPs.device
        N = frame_ids.numel()
        cameras, _, _ = self._cameras(N, device)
        poses, trans, d_cond, rendcond = self.dataset.get_grad_parameters(frame_ids, device)
        pix = torch.cat([col_inds.view(-1, 1), row_inds.view(-1, 1), torch.ones_like(col_inds.view(-1, 1))], dim=-1)
        with torch.no_grad():
            rays = cameras.view_rays(pix.float())
        defconds = [d_cond.detach(), [poses.detach(), trans.detach()]]
        tcolors = []
        cam_pos = cameras.cam_pos().detach()
        for rays_, ps_, bi_ in zip(torch.split(rays, chunk), torch.split(initTmpPs, chunk), torch.split(batch_inds, chunk)):
            ps_, check = utils.OptimizeSurfacePs(cam_pos, rays_.detach(), ps_.clone(), bi_, self.sdf, ratio,
                                                 self.deformer, defconds, dthreshold=1.e-4, athreshold=self.angThred,
                                                 w1=3.05, w2=1., times=30)
            if hasattr(self.sdf, "forward_fused"):
                tcolors.append(utils.shade_rays(self.sdf, self.deformer, self.netRender, ps_, rays_, defconds, bi_,
                                                ratio)[2])
                continue
            ps_ = ps_.detach().requires_grad_(True)   # user-supplied field modules: reference sequence
            sdfs = self.sdf(ps_, ratio)
            nx = torch.autograd.grad(sdfs, ps_, torch.ones_like(sdfs))[0]
            nx = nx / nx.norm(dim=1, keepdim=True)
            crays, defVs = utils.compute_cardinal_rays(self.deformer, ps_, rays_, defconds, bi_, ratio, 'test')
            with torch.no_grad():
                self.sdf(ps_, ratio)
                tcolors.append(utils.compute_netRender_color(self.netRender, ps_, defVs, nx, crays,
                                                             self.sdf.rendcond, None, ratio))
        tcolors = torch.clamp((torch.cat(tcolors, dim=0) / 2. + 0.5) * 255., min=0., max=255.)
        colors = torch.ones(N, H, W, 3, device=device) * 255.
        colors[batch_inds, row_inds, col_inds, :] = tcolors
        return colors

    # ---- ray part of forward(): network.py:509-644 ----------------------------------------------
    def forward_rays(self, datas, batch_inds, row_inds, col_inds, initTmpPs, ratio, frame_ids,
                     extra_points=None):
        """Loss terms that depend on rays: eikonal (grad_weight), colour, normal.  `extra_points`
        stands for the template vertices the reference adds to the eikonal sample set (:543)."""
        device = frame_ids.device
        conf = self.conf
        gtCs = datas['img'].to(device)
        N = gtCs.shape[0]
        cameras, H, W = self._cameras(N, device)
        pix = torch.cat([col_inds.view(-1, 1), row_inds.view(-1, 1), torch.ones_like(col_inds.view(-1, 1))], dim=-1)
        rays = cameras.view_rays(pix.float())
        poses, trans, d_cond, rendcond = self.dataset.get_grad_parameters(frame_ids, device)
        defconds = [d_cond, [poses, trans]]
        self.info = {}
        initTmpPs, check = utils.OptimizeSurfacePs(cameras.cam_pos().detach(), rays.detach(), initTmpPs, batch_inds,
                                                   self.sdf, ratio, self.deformer, defconds, dthreshold=5.e-5,
                                                   athreshold=self.angThred, w1=3.05, w2=1., times=10)
        self.info['rayInfo'] = (check.numel(), check.sum().item())
        self.TmpPs = None
        total_loss = torch.zeros((), device=device)
        # eikonal on jittered surface points + uniform samples (:543-549)
        base = initTmpPs if extra_points is None else torch.cat([initTmpPs, extra_points.detach()], dim=0)
        nonmnfld_pnts = utils.sample_points(base, 1.8, 0.01)
        nonmnfld_pnts.requires_grad_()
        pred = self.sdf(nonmnfld_pnts, ratio)
        grad = self.sdf.gradient(nonmnfld_pnts, pred)
        grad_loss = ((grad.norm(2, dim=-1) - 1) ** 2).mean()
        self.info['grad_loss'] = grad_loss.item()
        total_loss = total_loss + grad_loss * conf.get_float('grad_weight')
        self.info['color_loss'] = -1.0
        if self.info['rayInfo'][1] > 0:
            self.TmpPs = initTmpPs[check]
            self.TmpPs.requires_grad = True
            self.rays = rays[check]
            self.batch_inds = batch_inds[check]
            self.col_inds = col_inds[check]
            self.row_inds = row_inds[check]
            sdfs = self.sdf(self.TmpPs, ratio)
            nx = torch.autograd.grad(sdfs, self.TmpPs, torch.ones_like(sdfs), retain_graph=True, create_graph=True)[0]
            nx = nx / nx.norm(dim=1, keepdim=True)
            crays, defVs = utils.compute_cardinal_rays(self.deformer, self.TmpPs, self.rays, defconds,
                                                       self.batch_inds, ratio, 'train')
            if conf.get_float('color_weight') > 0.:
                colors = utils.compute_netRender_color(self.netRender, self.TmpPs, defVs, nx, crays, self.sdf.rendcond,
                                                       None, ratio)
                color_loss = (gtCs[self.batch_inds, self.row_inds, self.col_inds, :] - colors).abs().sum(1)
                color_loss = _scatter_mean(color_loss, self.batch_inds, N).mean()
                self.info['color_loss'] = color_loss.item()
                total_loss = total_loss + conf.get_float('color_weight') * color_loss
            if 'normal' in datas and 'normal_weight' in conf and conf.get_float('normal_weight') > 0.:
                if 'weighted_normal' in conf and conf.get_bool('weighted_normal'):
                    cnx, _ = utils.compute_deformed_normals(self.sdf, self.deformer, self.TmpPs, defconds,
                                                            self.batch_inds, ratio, 'test')
                    weights = torch.clamp((-self.rays * cnx.detach()).sum(1).detach(), max=1., min=0.) ** 2
                else:
                    weights = torch.ones(nx.shape[0], device=device)
                gtn = datas['normal'].to(device)[self.batch_inds, self.row_inds, self.col_inds, :]
                flip = torch.tensor([[-1., 0., 0.], [0., 1., 0.], [0., 0., -1.]], device=device)
                gtn = ((cameras.R[0] @ flip) @ gtn.view(-1, 3, 1)).view(-1, 3)
                gtnorms = gtn.norm(dim=1, keepdim=True)
                valid = (gtnorms > 0.0001)[..., 0]
                gtn = torch.where(valid.view(-1, 1), gtn / gtnorms.clamp(min=1e-12), gtn)
                ds = self.deformer(self.TmpPs, defconds, self.batch_inds, ratio=ratio)
                J = utils.compute_Jacobian(self.TmpPs, ds, True, True)
                gtn = (J.transpose(-2, -1) @ gtn.view(-1, 3, 1)).view(-1, 3)
                normal_loss = (gtn - nx).norm(2, dim=1) * weights
                normal_loss = _scatter_mean(normal_loss[valid], self.batch_inds[valid], N).mean()
                self.info['normal_loss'] = normal_loss.item()
                total_loss = total_loss + conf.get_float('normal_weight') * normal_loss
        self.forward_time += 1
        return total_loss

    def forward(self, datas, sample_pix, ratio, frame_ids, root=None, **kwargs):
        if self.raster_seed is None:
            raise RuntimeError("OptimNetwork.forward needs the raster seed of network.py:485-505 "
                               "(pytorch3d MeshRasterizer / PointsRasterizer): set `raster_seed` to a "
                               "callable or call forward_rays() with your own seed")
        device = frame_ids.device
        if self.TmpVs is None or self.Tmpfs is None or self.forward_time % self.remesh_intersect == 0:
            self.TmpVs, self.Tmpfs = self.discretizeSDF(ratio, None, -self.sdfShrinkRadius)
            if self.TmpVs.shape[0] == 0:
                print('tmp sdf vanished...')
                assert False
        poses, trans, d_cond, _ = self.dataset.get_grad_parameters(frame_ids, device)
        seed = self.raster_seed(frame_ids, self.TmpVs, self.Tmpfs, [d_cond, [poses, trans]], ratio)
        bi, ri, ci, ps = seed['batch_inds'], seed['row_inds'], seed['col_inds'], seed['initTmpPs']
        gtMs = datas['mask'].to(device)
        sel = gtMs[bi, ri, ci] > 0.
        bi, ri, ci, ps = bi[sel], ri[sel], ci[sel], ps[sel]
        N = frame_ids.numel()
        sample_pix = self.conf.get_int('sample_pix_num') if 'sample_pix_num' in self.conf else sample_pix
        if bi.shape[0] > sample_pix * N:
            sel = (torch.rand(bi.shape[0]) < float(sample_pix * N) / float(bi.shape[0])).to(device)
            bi, ri, ci, ps = bi[sel], ri[sel], ci[sel], ps[sel]
        TmpVnum = self.TmpVs.shape[0]
        extra = self.TmpVs[(torch.rand(TmpVnum) < 4096. / float(TmpVnum)).to(device)]
        return self.forward_rays(datas, bi, ri, ci, ps, ratio, frame_ids, extra_points=extra)

    # ---- network.py:702-814 ---------------------------------------------------------------------
    def propagateTmpPsGrad(self, frame_ids, ratio):
        if self.TmpPs is None or self.TmpPs.grad is None:
            self.info['invInfo'] = (-1, -1)
            return
        device = self.TmpPs.device
        poses, trans, d_cond, _ = self.dataset.get_grad_parameters(frame_ids, device)
        defconds = [d_cond, [poses, trans]]
        cameras, _, _ = self._cameras(frame_ids.numel(), device)
        grad_l_p = self.TmpPs.grad
        if self.rays.requires_grad:
            pix = torch.cat([self.col_inds.view(-1, 1), self.row_inds.view(-1, 1),
                             torch.ones_like(self.col_inds.view(-1, 1))], dim=-1)
            v = cameras.view_rays(pix.float())
        else:
            v = self.rays.detach()
        c = cameras.cam_pos()
        p = self.TmpPs
        fusable = hasattr(self.sdf, "forward_fused") and hasattr(self.deformer, "forward_fused") \
            and self.deformer._fusable()
        if fusable:
            # grad f and dD/dp at p from the fused forward-mode kernels: no graph, no 1+3 VJP passes
            with torch.no_grad():
                _, grad_f_p, _ = self.sdf.forward_fused(p.detach(), ratio, want_grad=True, want_feat=False)
                _, grad_d_p, _ = self.deformer.forward_fused(p.detach(), [c_.detach() if torch.is_tensor(c_) else
                                                                          [t.detach() for t in c_] for c_ in defconds],
                                                             self.batch_inds, ratio, want_jac=True)
        else:
            f = self.sdf(p, ratio)
            grad_f_p = torch.autograd.grad(f, p, torch.ones_like(f), retain_graph=False)[0]
            d = self.deformer(p, defconds, self.batch_inds, ratio=ratio)
            grad_d_p = utils.compute_Jacobian(p, d, False, False)
        v_cross = _cross_matrix(v.detach())
        a1 = v_cross.matmul(grad_d_p)
        b = torch.cat([grad_f_p.view(-1, 1, 3), a1], dim=1)           # [P,4,3]
        btb = b.permute(0, 2, 1).matmul(b)
        btb_inv, check = Fast3x3Minv(btb.contiguous())
        self.info['invInfo'] = (check.numel(), check.sum().item())
        rhs_1 = grad_l_p.view(-1, 1, 3).matmul(btb_inv.matmul(b.permute(0, 2, 1)))   # [P,1,4]
        loss = 0.
        # theta: VJP of the sdf with cotangent -rhs[...,0]
        params = [q for q in self.sdf.parameters() if q.requires_grad]
        grads = torch.autograd.grad(self.sdf(p, ratio), params, -rhs_1[:, :, 0])
        for q, g in zip(params, grads):
            loss = loss + (q * g).sum()
        # phi, latent codes, pose: VJP of the deformer with cotangent rhs[...,1:] (-[v]x)
        opt_defconds = []
        for dc in defconds:
            for t in (dc if isinstance(dc, list) else [dc]):
                if t.requires_grad:
                    opt_defconds.append(t)
        params = [q for q in self.deformer.parameters() if q.requires_grad]
        d = self.deformer(p, defconds, self.batch_inds, ratio=ratio)
        temp = (rhs_1[:, :, -3:].matmul(-v_cross)).view(-1, 3)
        grads = torch.autograd.grad(d, params + opt_defconds, temp)
        for q, g in zip(params + opt_defconds, grads):
            loss = loss + (q * g).sum()
        if v.requires_grad:
            dc_cross = _cross_matrix(d.detach() - c.detach().view(1, 3))
            loss = loss + (v * rhs_1[:, :, -3:].matmul(dc_cross).view(-1, 3)).sum()
        if c.requires_grad:
            loss = loss + (c * (-temp.sum(0))).sum()
        loss.backward()
