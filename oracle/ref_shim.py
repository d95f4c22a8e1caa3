"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (``/root/reference``) on CPU by registering
placeholder modules for the packages that are absent from this image
(pytorch3d, torch_scatter, pyhocon, trimesh, openmesh, the compiled CUDA ops).
Only usable inside the build container (the GPU box has no /root/reference);
it exists to (a) validate ``oracle/oracle.py`` against the real reference and
(b) generate the committed golden vectors under ``tests/golden/`` (see
``oracle/make_golden.py``).

CPU stand-ins supplied for the compiled ops follow the reference sources:
  * ``GridSamplerMine.forward/backward`` -> ``F.grid_sample(bilinear, border,
    align_corners=False)`` which is the call the reference author replaced
    (model/Deformer.py:208-211).
  * ``FastMinv.Fast3x3Minv`` -> cofactor inverse with the |det|<1e-4 mask
    (FastMinv/Matrix3x3InvKernels.cu:22-61).
  * ``torch_scatter.scatter`` -> mean/min reductions (call sites
    model/network.py:617,637, utils/FindSurfacePs.py:15).
"""
import sys
import types
import importlib

REF_ROOT = "/root/reference"


class _Anything:
    """Attribute sink: any attribute / call returns another sink (for unused imports)."""

    def __init__(self, name="stub"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __call__(self, *a, **k):
        return _Anything(self._n + "()")

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package

    def __getattr__(k, _n=name):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(_n + "." + k)

    m.__getattr__ = __getattr__
    return m


def _install_stubs():
    import torch
    import torch.nn.functional as F

    names = [
        "pytorch3d", "pytorch3d.structures", "pytorch3d.loss", "pytorch3d.io",
        "pytorch3d.renderer", "pytorch3d.renderer.cameras", "pytorch3d.renderer.mesh",
        "pytorch3d.renderer.mesh.renderer", "pytorch3d.renderer.points",
        "pytorch3d.renderer.utils", "pytorch3d.transforms", "pytorch3d.ops",
        "pytorch3d.renderer.mesh.rasterizer", "pytorch3d.renderer.points.rasterizer",
        "pytorch3d.common", "pytorch3d.common.types", "pytorch3d.renderer.mesh.shader",
        "pytorch3d.renderer.blending", "pytorch3d.renderer.lighting",
        "pyhocon", "trimesh", "openmesh", "cv2", "MCGpu", "interp2x_boundary3d",
        "interp2x_boundary2d", "h5py", "skimage",
    ]
    for n in names:
        if n not in sys.modules:
            try:
                if n == "cv2":
                    importlib.import_module(n)
                    continue
            except Exception:
                pass
            sys.modules[n] = _stub_module(n)

    # ---- torch_scatter -------------------------------------------------
    ts = types.ModuleType("torch_scatter")

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
        assert dim == 0 or src.dim() == 1
        if reduce == "mean":
            n = dim_size if dim_size is not None else int(index.max().item()) + 1
            s = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            s = s.index_add(0, index, src)
            c = torch.zeros(n, dtype=src.dtype, device=src.device).index_add(
                0, index, torch.ones_like(index, dtype=src.dtype))
            c = c.clamp(min=1)
            return s / c.view((-1,) + (1,) * (src.dim() - 1))
        if reduce == "min":
            assert out is not None
            return out.scatter_reduce(0, index, src, reduce="amin", include_self=True)
        if reduce == "sum":
            n = dim_size if dim_size is not None else int(index.max().item()) + 1
            s = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            return s.index_add(0, index, src)
        raise NotImplementedError(reduce)

    ts.scatter = scatter
    sys.modules["torch_scatter"] = ts

    # ---- FastMinv -------------------------------------------------------
    fm = types.ModuleType("FastMinv")

    def Fast3x3Minv(ms):
        m = ms
        c00 = m[:, 1, 1] * m[:, 2, 2] - m[:, 1, 2] * m[:, 2, 1]
        c01 = -m[:, 1, 0] * m[:, 2, 2] + m[:, 1, 2] * m[:, 2, 0]
        c02 = m[:, 1, 0] * m[:, 2, 1] - m[:, 1, 1] * m[:, 2, 0]
        c10 = -m[:, 0, 1] * m[:, 2, 2] + m[:, 0, 2] * m[:, 2, 1]
        c11 = m[:, 0, 0] * m[:, 2, 2] - m[:, 0, 2] * m[:, 2, 0]
        c12 = -m[:, 0, 0] * m[:, 2, 1] + m[:, 0, 1] * m[:, 2, 0]
        c20 = m[:, 0, 1] * m[:, 1, 2] - m[:, 0, 2] * m[:, 1, 1]
        c21 = -m[:, 0, 0] * m[:, 1, 2] + m[:, 0, 2] * m[:, 1, 0]
        c22 = m[:, 0, 0] * m[:, 1, 1] - m[:, 0, 1] * m[:, 1, 0]
        det = m[:, 0, 0] * c00 + m[:, 0, 1] * c01 + m[:, 0, 2] * c02
        ok = ~(det.abs() < 0.0001)
        adj = torch.stack([c00, c10, c20, c01, c11, c21, c02, c12, c22], 1).view(-1, 3, 3)
        inv = adj / torch.where(ok, det, torch.ones_like(det)).view(-1, 1, 1)
        inv = torch.where(ok.view(-1, 1, 1), inv, torch.zeros_like(inv))
        return [inv.detach(), ok]

    def Fast3x3Minv_backward(grads, invs):
        ct = invs.transpose(1, 2)
        return -(ct @ grads @ ct)

    fm.Fast3x3Minv = Fast3x3Minv
    fm.Fast3x3Minv_backward = Fast3x3Minv_backward
    sys.modules["FastMinv"] = fm

    # ---- GridSamplerMine: forward / backward / dbackward from ONE differentiable torch restatement of the
    #      sampler (trilinear, border padding, align_corners=False: MCAcc/cuda/GridSamplerMineKernel.cu:160-309),
    #      so first and second order are consistent with each other on CPU ----
    gs = types.ModuleType("GridSamplerMine")

    def _trilinear(inp, grid):
        n, c, D, H, W = inp.shape
        assert n == 1
        g = grid.reshape(-1, 3)
        size = torch.tensor([W, H, D], dtype=g.dtype)
        ix = ((g + 1.0) * size - 1.0) / 2.0
        ix = torch.minimum(torch.maximum(ix, torch.zeros(3, dtype=g.dtype)), size - 1.0)     # border clip
        i0 = torch.floor(ix.detach())
        fr = ix - i0
        i0 = i0.long()
        vol = inp[0].reshape(c, -1)
        out = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi = (i0[:, 0] + dx).clamp(max=W - 1)
                    yi = (i0[:, 1] + dy).clamp(max=H - 1)
                    zi = (i0[:, 2] + dz).clamp(max=D - 1)
                    w = (fr[:, 0] if dx else 1 - fr[:, 0]) * (fr[:, 1] if dy else 1 - fr[:, 1]) * \
                        (fr[:, 2] if dz else 1 - fr[:, 2])
                    out = out + vol[:, (zi * H + yi) * W + xi] * w.unsqueeze(0)
        return out.reshape((1, c) + tuple(grid.shape[1:4]))

    def gs_forward(inp, grid, interp=0, pad=1):
        with torch.no_grad():
            return _trilinear(inp, grid)

    def gs_backward(inp, grid, gout, interp=0, pad=1):
        with torch.enable_grad():
            i = inp.detach().requires_grad_(True)
            g = grid.detach().requires_grad_(True)
            gi, gg = torch.autograd.grad(_trilinear(i, g), [i, g], gout)
        return gi, gg

    def gs_dbackward(gg_inp, gg_grid, inp, grid, gout, interp=0, pad=1):
        with torch.enable_grad():
            i = inp.detach().requires_grad_(True)
            g = grid.detach().requires_grad_(True)
            go = gout.detach().requires_grad_(True)
            gi, gg = torch.autograd.grad(_trilinear(i, g), [i, g], go, create_graph=True)
            s = (gi * gg_inp).sum() + (gg * gg_grid).sum()
            a, b, c = torch.autograd.grad(s, [i, g, go], allow_unused=True)
        z = lambda t, like: torch.zeros_like(like) if t is None else t
        return z(a, inp), z(b, grid), z(c, gout)

    gs.forward = gs_forward
    gs.backward = gs_backward
    gs.dbackward = gs_dbackward
    sys.modules["GridSamplerMine"] = gs


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's hot-path modules imported on CPU."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import MCAcc  # noqa
    import utils  # noqa
    import model.network as network
    import model.Deformer as Deformer
    import model.RenderNet as RenderNet
    import model.Embedder as Embedder
    import MCAcc.seg3d_lossless as seg3d
    import smpl_pytorch.util as smpl_util
    _loaded.update(dict(MCAcc=MCAcc, utils=utils, network=network, Deformer=Deformer,
                        RenderNet=RenderNet, Embedder=Embedder, seg3d=seg3d, smpl_util=smpl_util))
    return types.SimpleNamespace(**_loaded)


if __name__ == "__main__":
    ref = load_reference()
    import torch
    torch.manual_seed(0)
    net = ref.network.getTmpSdf("cpu", 6)
    x = torch.randn(8, 3) * 0.5
    print(net(x, 1.0).view(-1))
