"""Camera of the hot path (reference: model/CameraMine.py:129-170).

Only the three methods the per-frame step uses are mirrored -- view_rays, cam_pos, project,
angThreshold -- on a plain class with the reference's attribute names (focal_length,
principal_point, R, T, image_size).  The pytorch3d `CamerasBase` plumbing the reference adds so
that pytorch3d's rasterisers accept the object (CameraMine.py:15-127) is out of scope; when
pytorch3d is installed the reference's own class can be used instead: this module's functions
only rely on the attributes above.  Everything stays differentiable w.r.t. the camera tensors."""
import numpy as np
import torch


class RectifiedPerspectiveCameras:
    def __init__(self, focal_length, principal_point, R, T, image_size=None, device=None):
        self.focal_length = focal_length
        self.principal_point = principal_point
        self.R = R
        self.T = T
        if image_size is not None and not torch.is_tensor(image_size):
            image_size = torch.tensor(image_size)
        self.image_size = image_size
        if device is not None:
            self.to(device)

    def to(self, device):
        for k in ("focal_length", "principal_point", "R", "T", "image_size"):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def view_rays(self, ps, cam_id=0):
        f, c = self.focal_length[cam_id], self.principal_point[cam_id]
        rays = torch.stack([-ps[:, 0] / f[0] + ps[:, 2] * c[0] / f[0],
                            -ps[:, 1] / f[1] + ps[:, 2] * c[1] / f[1], ps[:, 2]], dim=1)
        rays = rays / torch.norm(rays, p=2, dim=1, keepdim=True)
        # rays R^T as a broadcast multiply + sum: stays differentiable w.r.t. R, launches no GEMM
        return (rays.unsqueeze(1) * self.R[cam_id].unsqueeze(0)).sum(2)

    def project(self, ps, cam_id=0):
        ps = (ps.unsqueeze(2) * self.R[cam_id].unsqueeze(0)).sum(1) + self.T[cam_id].view(1, 3)
        x = self.principal_point[cam_id, 0] - ps[:, 0] * self.focal_length[cam_id, 0] / ps[:, 2]
        y = self.principal_point[cam_id, 1] - ps[:, 1] * self.focal_length[cam_id, 1] / ps[:, 2]
        return torch.cat([x.view(-1, 1), y.view(-1, 1)], dim=1)

    def cam_pos(self, cam_id=0):
        return -(self.R[cam_id] * self.T[cam_id].view(1, 3)).sum(1)

    def angThreshold(self, pixoffset=0.4, cam_id=0):
        H = self.image_size[cam_id, 1].item()
        W = self.image_size[cam_id, 0].item()
        cx, cy = self.principal_point[cam_id, 0].item(), self.principal_point[cam_id, 1].item()
        fx, fy = self.focal_length[cam_id, 0].item(), self.focal_length[cam_id, 1].item()

        def ang(a, b):
            r1, r2 = torch.tensor(a), torch.tensor(b)
            return torch.arcsin(torch.linalg.cross(r1, r2).norm() / (r1.norm() * r2.norm())) / np.pi * 180.

        th = ang([(W - cx) / fx, 0., 1.], [(W + pixoffset - cx) / fx, 0., 1.])
        th = torch.min(th, ang([-cx / fx, 0., 1.], [(pixoffset - cx) / fx, 0., 1.]))
        th = torch.min(th, ang([0., (H - cy) / fy, 1.], [0., (H + pixoffset - cy) / fy, 1.]))
        th = torch.min(th, ang([0., -cy / fy, 1.], [0., (pixoffset - cy) / fy, 1.]))
        return th.item()


class PointsRendererWithFrags(torch.nn.Module):
    """Soft point-cloud silhouette renderer that also returns the rasteriser's fragments
    (model/CameraMine.py:283-320): weight 1 - d^2/r^2 per (pixel, point) pair, composited by the given
    compositor.  Works with any rasteriser / compositor pair following pytorch3d's protocol
    (`rasterizer(point_clouds) -> fragments(idx, dists)`, `raster_settings.radius`)."""

    def __init__(self, rasterizer, compositor):
        super().__init__()
        self.rasterizer = rasterizer
        self.compositor = compositor

    def to(self, device):
        self.rasterizer = self.rasterizer.to(device)
        self.compositor = self.compositor.to(device)
        return self

    def forward(self, point_clouds, **kwargs):
        frags = self.rasterizer(point_clouds, **kwargs)
        r = self.rasterizer.raster_settings.radius
        weights = 1 - frags.dists.permute(0, 3, 1, 2) / (r * r)
        images = self.compositor(frags.idx.long().permute(0, 3, 1, 2), weights,
                                 point_clouds.features_packed().permute(1, 0), **kwargs)
        return images.permute(0, 2, 3, 1), frags
