"""Built-in silhouette rasteriser for the ray seed.

`MeshRasterizer` here plays the role pytorch3d's MeshRasterizer plays at model/network.py:864-880 (the object
behind `optNet.maskRender.rasterizer`): it owns `.cameras` and `.raster_settings` and turns the deformed template
into Fragments(pix_to_face, bary_coords, zbuf) -- on one CUDA kernel pair (csrc/raster.cu) instead of the
binned pytorch3d pipeline, and on the camera convention of RectifiedPerspectiveCameras.project directly
(pixel centres at integer (col, row)), so no NDC round trip.  `SilhouetteRenderer` is the minimal
MeshRendererWithFragments: `renderer(verts [N,V,3], faces) -> (hard silhouette [N,H,W,4], fragments)`."""
import types

import torch

from selfreconcode_b200 import ops


class RasterSettings:
    def __init__(self, image_size, blur_radius=0., faces_per_pixel=1, perspective_correct=True,
                 clip_barycentric_coords=False, cull_backfaces=False, bin_size=None):
        self.image_size = tuple(image_size)
        self.blur_radius = blur_radius
        self.faces_per_pixel = faces_per_pixel
        self.perspective_correct = perspective_correct
        self.clip_barycentric_coords = clip_barycentric_coords
        self.cull_backfaces = cull_backfaces
        self.bin_size = bin_size
        if blur_radius != 0. or faces_per_pixel != 1 or cull_backfaces or clip_barycentric_coords:
            raise NotImplementedError("the built-in rasteriser implements the settings the optimisation step uses: "
                                      "blur 0, one face per pixel, no culling, unclipped barycentrics")


class MeshRasterizer:
    def __init__(self, cameras, raster_settings):
        self.cameras = cameras
        self.raster_settings = raster_settings

    def to(self, device):
        if hasattr(self.cameras, "to"):
            self.cameras = self.cameras.to(device)
        return self

    def screen_vertices(self, verts, cameras=None):
        """[N,V,3] world -> (col, row, depth) with frame n seen by camera n (CameraMine.py:138-142)."""
        cam = cameras if cameras is not None else self.cameras
        N = verts.shape[0]
        R, T = cam.R[:N], cam.T[:N]
        pc = (verts.unsqueeze(3) * R.unsqueeze(1)).sum(2) + T.view(N, 1, 3)          # p R + T, no GEMM launch
        f, c = cam.focal_length[:N], cam.principal_point[:N]
        x = c[:, 0:1] - pc[..., 0] * f[:, 0:1] / pc[..., 2]
        y = c[:, 1:2] - pc[..., 1] * f[:, 1:2] / pc[..., 2]
        return torch.stack([x, y, pc[..., 2]], dim=-1)

    def __call__(self, verts, faces, cameras=None):
        H, W = self.raster_settings.image_size
        with torch.no_grad():
            vs = self.screen_vertices(verts, cameras)
            p2f, bary, zbuf = ops.raster_mesh(vs, faces, H, W)
        return types.SimpleNamespace(pix_to_face=p2f, bary_coords=bary, zbuf=zbuf,
                                     dists=torch.zeros_like(zbuf))


class SilhouetteRenderer:
    """`images, fragments = renderer(verts, faces)`; images[..., 3] is the hard coverage mask."""

    def __init__(self, rasterizer):
        self.rasterizer = rasterizer

    def to(self, device):
        self.rasterizer.to(device)
        return self

    def __call__(self, verts, faces, cameras=None, **kwargs):
        frags = self.rasterizer(verts, faces, cameras)
        cover = (frags.pix_to_face >= 0).float()
        return torch.cat([cover.expand(-1, -1, -1, 3), cover], dim=-1), frags
