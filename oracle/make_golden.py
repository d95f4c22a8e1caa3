"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run once in the build container (needs /root/reference; the GPU box does not have it):
    python oracle/make_golden.py
Every array written here comes out of the reference's own Python (imported on CPU through
oracle/ref_shim.py) on seeded inputs; tests replay the same inputs through oracle/oracle.py
(CPU, `-m "not gpu"`) and through the CUDA path (`-m gpu`).

Large parameter sets are not stored: MLPTranslator is hard-wired to 512-wide layers
(model/Deformer.py:25), so those fixtures store the torch seed plus a checksum of the
parameters; the tests rebuild the parameters with the same constructor sequence and verify the
checksum before using them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
RATIO = {"sdfRatio": 1.0, "deformerRatio": 0.8, "renderRatio": 1.0}

SMPL_PARENTS = np.array([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])


def checksum(module):
    s = 0.0
    a = 0.0
    for _, p in sorted(module.state_dict().items()):
        s += float(p.double().sum())
        a += float(p.double().abs().sum())
    return np.array([s, a], dtype=np.float64)


def sd_numpy(module, prefix=""):
    return {prefix + k.replace(".", "__"): v.detach().numpy() for k, v in module.state_dict().items()}


def perturb(module, scale, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.add_(scale * torch.randn(p.shape, generator=g))


def synth_ws(D, H, W, bmin, bmax, Js, sigma=0.25):
    xs = (torch.arange(W).float() + 0.5) / W * (bmax[0] - bmin[0]) + bmin[0]
    ys = (torch.arange(H).float() + 0.5) / H * (bmax[1] - bmin[1]) + bmin[1]
    zs = (torch.arange(D).float() + 0.5) / D * (bmax[2] - bmin[2]) + bmin[2]
    logits = torch.empty(24, D, H, W)
    for j in range(24):
        d2 = ((zs - Js[j, 2]) ** 2).view(D, 1, 1) + ((ys - Js[j, 1]) ** 2).view(1, H, 1) + \
             ((xs - Js[j, 0]) ** 2).view(1, 1, W)
        logits[j] = -d2 / (2 * sigma * sigma)
    return torch.softmax(logits, 0).unsqueeze(0).contiguous()


def grad_digest(grad, idx):
    """[norm, dot with a seeded gaussian, <=128 strided samples] of a gradient tensor."""
    flat = grad.detach().double().reshape(-1)
    r = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(9000 + idx), dtype=torch.float64)
    stride = max(1, flat.numel() // 128)
    return np.concatenate([[flat.norm().item(), (flat * r).sum().item()], flat[::stride][:128].numpy()])


def pack_grid(grid, queried):
    """Small fixture for a coarse-to-fine grid: the queried mask, the exact values at queried
    voxels (z,y,x order), the sign pattern of the whole grid (what marching cubes consumes) and
    a random sample of interpolated-only voxels (those are only reproducible to ~1e-6)."""
    g = grid.numpy()
    q = queried.numpy()
    rs = np.random.RandomState(5)
    rest = np.flatnonzero(~q.reshape(-1))
    samp = rs.choice(rest, size=min(4000, rest.size), replace=False)
    return dict(shape=np.array(g.shape), queried=np.packbits(q), values_at_queried=g[q],
                sign=np.packbits(g > 0.0), interp_idx=samp.astype(np.int64),
                interp_val=g.reshape(-1)[samp])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    net_mod, def_mod, rnd_mod, utils = ref.network, ref.Deformer, ref.RenderNet, ref.utils

    # ---------------------------------------------------------------- embedder / annealing
    x = torch.linspace(-1.3, 1.7, 30).view(10, 3)
    emb, odim = ref.Embedder.get_embedder(6)
    out = {"x": x.numpy(), "embed6": emb(x).numpy()}
    for r in (0.05, 0.3, 0.55, 1.0):
        ws = utils.annealing_weights(6, r)
        out["aw_%g" % r] = np.array(ws, dtype=np.float64)
        out["embed6_r%g" % r] = emb(x, ws).numpy()
    np.savez(os.path.join(OUT, "embedder.npz"), **out)

    # ---------------------------------------------------------------- small SDF (params stored)
    torch.manual_seed(11)
    sdf = net_mod.ImplicitNetwork(16, 3, 1, [64, 64, 64, 64], geometric_init=True, bias=0.6,
                                  skip_in=[2], weight_norm=True, multires=6)
    perturb(sdf, 2e-2, 12)
    g = torch.Generator().manual_seed(13)
    pts = (torch.rand(300, 3, generator=g) - 0.5) * 1.6
    out = sd_numpy(sdf, "p_")
    for r in (1.0, 0.4):
        p = pts.clone().requires_grad_(True)
        s = sdf(p, r)
        (gr,) = torch.autograd.grad(s, p, torch.ones_like(s))
        out["sdf_r%g" % r] = s.detach().numpy()
        out["grad_r%g" % r] = gr.numpy()
        out["feat_r%g" % r] = sdf.rendcond.detach().numpy()
    out["pts"] = pts.numpy()
    np.savez(os.path.join(OUT, "sdf_small.npz"), **out)

    # ---------------------------------------------------------------- full-size SDF (seeded)
    torch.manual_seed(0)
    # bias 0.78 + 3e-3 noise: a lumpy sphere of mean radius ~0.6 (with the reference default
    # bias 0.6 the PE/softplus network's zero set sits at r~0.34, and 1e-2 noise on every weight
    # removes the zero set altogether -- "tmp sdf vanished", network.py:466-468)
    sdf_full = net_mod.getTmpSdf("cpu", 6, bias=0.78)
    perturb(sdf_full, 3e-3, 1000)
    g = torch.Generator().manual_seed(21)
    pts = (torch.rand(96, 3, generator=g) - 0.5) * 1.5
    p = pts.clone().requires_grad_(True)
    s = sdf_full(p, 1.0)
    (gr,) = torch.autograd.grad(s, p, torch.ones_like(s))
    np.savez(os.path.join(OUT, "sdf_full.npz"), seed=0, perturb_seed=1000, perturb=3e-3, bias=0.78,
             checksum=checksum(sdf_full), pts=pts.numpy(), sdf=s.detach().numpy(),
             grad=gr.numpy(), feat=sdf_full.rendcond.detach().numpy())

    # ---------------------------------------------------------------- translator + LBS (seeded)
    torch.manual_seed(1)
    tr = def_mod.MLPTranslator(128, 6)
    perturb(tr, 1e-2, 1001)
    bmin, bmax = [-0.9, -1.3, -0.5], [0.9, 0.9, 0.5]
    gJ = torch.Generator().manual_seed(31)
    Js = (torch.rand(24, 3, generator=gJ) - 0.5) * torch.tensor([1.4, 1.8, 0.6])
    D, H, W = 7, 13, 9
    ws = synth_ws(D, H, W, bmin, bmax, Js)
    apose = utils.smpl_tmp_Apose(1)
    sk = def_mod.LBSkinner(ws, bmin, bmax, Js, SMPL_PARENTS, init_pose=apose)
    comp = def_mod.CompositeDeformer([tr, sk])
    N = 3
    g = torch.Generator().manual_seed(32)
    poses = 0.2 * torch.randn(N, 24, 3, generator=g)
    trans = 0.05 * torch.randn(N, 3, generator=g)
    dcond = 0.1 * torch.randn(N, 128, generator=g)
    P = 120
    pts = (torch.rand(P, 3, generator=g) - 0.5) * torch.tensor([2.2, 2.6, 1.3])  # some outside box
    bi = torch.randint(0, N, (P,), generator=g)
    p = pts.clone().requires_grad_(True)
    d = comp(p, [dcond, [poses, trans]], bi, ratio=RATIO)
    J = utils.compute_Jacobian(p, d, True, False)
    off = tr.offset.detach()
    # LBS alone + posed skeleton + bone transforms
    d_lbs = sk(pts, [poses, trans], bi)
    posed = sk.posedSkeleton([poses, trans])
    np.savez(os.path.join(OUT, "deform.npz"), seed=1, perturb_seed=1001, perturb=1e-2,
             checksum=checksum(tr), Js=Js.numpy(), ws=ws.numpy(), bmin=np.array(bmin, np.float32),
             bmax=np.array(bmax, np.float32), apose=apose, init_pose_inv=sk.init_pose.numpy(),
             poses=poses.numpy(), trans=trans.numpy(), dcond=dcond.numpy(), pts=pts.numpy(),
             batch_inds=bi.numpy(), d=d.detach().numpy(), jac=J.numpy(), offset=off.numpy(),
             d_lbs=d_lbs.numpy(), posed=posed.numpy(), def_ratio=RATIO["deformerRatio"])

    # cardinal rays through the reference helper (uses the shim's Fast3x3Minv stand-in)
    rays = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1)
    p = pts.clone().requires_grad_(True)
    crays, ds = utils.compute_cardinal_rays(comp, p, rays, [dcond, [poses, trans]], bi, RATIO, 'test')
    np.savez(os.path.join(OUT, "cardinal.npz"), rays=rays.numpy(), crays=crays.detach().numpy(),
             ds=ds.detach().numpy())

    # ---------------------------------------------------------------- render net (seeded)
    torch.manual_seed(2)
    rn = rnd_mod.RenderingNetwork_view_norm(256, 'idr', 9, 3, [512] * 4, weight_norm=True,
                                            multires_v=4, multires_n=0)
    g = torch.Generator().manual_seed(41)
    Pr = 80
    rp = (torch.rand(Pr, 3, generator=g) - 0.5)
    rn_ = torch.nn.functional.normalize(torch.randn(Pr, 3, generator=g), dim=1)
    rv = torch.nn.functional.normalize(torch.randn(Pr, 3, generator=g), dim=1)
    rf = 0.3 * torch.randn(Pr, 256, generator=g)
    col = rn(rp, rn_, rv, rf, RATIO)
    np.savez(os.path.join(OUT, "render.npz"), seed=2, checksum=checksum(rn), pts=rp.numpy(),
             normals=rn_.numpy(), views=rv.numpy(), feat=rf.numpy(), rgb=col.detach().numpy())

    # deformed normals through the reference helper (utils/utils.py:132-153), full-size SDF + composite deformer
    p = pts.clone().requires_grad_(True)
    dn, dds = utils.compute_deformed_normals(sdf_full, comp, p, [dcond, [poses, trans]], bi, RATIO, 'test')
    np.savez(os.path.join(OUT, "normals.npz"), normals=dn.detach().numpy(), ds=dds.detach().numpy())

    # ---------------------------------------------------------------- OptimizeSurfacePs
    # full-size sdf + composite deformer from above, a handful of rays aimed through D(p*)
    g = torch.Generator().manual_seed(51)
    Pn = 48
    dirs = torch.nn.functional.normalize(torch.randn(Pn, 3, generator=g), dim=1)
    dirs[:, 2] = -dirs[:, 2].abs()
    lo, hi = torch.full((Pn,), 0.05), torch.full((Pn,), 1.5)
    with torch.no_grad():
        for _ in range(40):  # bisection onto the zero level set along each direction
            mid = (lo + hi) / 2
            fm = sdf_full(mid.view(-1, 1) * dirs, RATIO).view(-1)
            lo, hi = torch.where(fm < 0, mid, lo), torch.where(fm < 0, hi, mid)
    pstar = ((lo + hi) / 2).view(-1, 1) * dirs
    bi2 = torch.randint(0, N, (Pn,), generator=g)
    cam_pos = torch.tensor([0.0, 0.0, -2.5])
    with torch.no_grad():
        dstar = comp(pstar, [dcond, [poses, trans]], bi2, ratio=RATIO)
    v = torch.nn.functional.normalize(dstar - cam_pos.view(1, 3), dim=1)
    start = pstar + 5e-3 * torch.randn(Pn, 3, generator=g)
    res = {}
    for name, (dth, times) in {"train": (5e-5, 10), "infer": (1e-4, 30)}.items():
        out_p, conv = utils.OptimizeSurfacePs(cam_pos, v, start.clone(), bi2, sdf_full, RATIO, comp,
                                              [dcond, [poses, trans]], dthreshold=dth,
                                              athreshold=0.0112, w1=3.05, w2=1., times=times)
        res["pts_" + name] = out_p.numpy()
        res["conv_" + name] = conv.numpy()
        with torch.no_grad():
            res["f_" + name] = sdf_full(out_p, RATIO).view(-1).numpy()
    np.savez(os.path.join(OUT, "trace.npz"), rays=v.numpy(), start=start.numpy(),
             batch_inds=bi2.numpy(), cam_pos=cam_pos.numpy(), athreshold=0.0112, **res)

    # ---------------------------------------------------------------- propagateTmpPsGrad
    # OptimNetwork.propagateTmpPsGrad (network.py:702-814) run by the reference itself on fake
    # dataset / renderer holders; the camera is the reference's RectifiedPerspectiveCameras with the
    # pytorch3d base-class constructor bypassed (its view_rays / cam_pos / angThreshold are used).
    import model.CameraMine as ref_cam

    class GoldCam(ref_cam.RectifiedPerspectiveCameras):
        def __init__(self, focal_length, principal_point, R, T, image_size=None, **kw):
            self.focal_length, self.principal_point, self.R, self.T = focal_length, principal_point, R, T
            self.image_size = torch.tensor(image_size)

        def to(self, device):
            return self

    net_mod.RectifiedPerspectiveCameras = GoldCam
    g = torch.Generator().manual_seed(71)
    Hh = Ww = 64
    focals0 = torch.tensor([[150.0, 152.0]]).repeat(N, 1)
    pps0 = torch.tensor([[31.5, 32.5]]).repeat(N, 1)
    Rs0 = torch.tensor([[[-1., 0., 0.], [0., 1., 0.], [0., 0., -1.]]]).repeat(N, 1, 1)
    Ts0 = torch.tensor([[0.02, -0.01, 2.5]]).repeat(N, 1)
    cam0 = GoldCam(focals0, pps0, Rs0, Ts0, image_size=[(Ww, Hh)])
    Pg = 40
    with torch.no_grad():
        dg = comp(pstar[:Pg], [dcond, [poses, trans]], bi2[:Pg], ratio=RATIO)
        pix = cam0.project(dg)
    col = pix[:, 0].round().long()
    row = pix[:, 1].round().long()
    tmpps = pstar[:Pg] + 2e-3 * torch.randn(Pg, 3, generator=g)
    gl = torch.randn(Pg, 3, generator=g)
    out = dict(cam_project_in=dg.numpy(), cam_project=pix.numpy(), cam_pos=cam0.cam_pos().numpy(),
               col=col.numpy(), row=row.numpy(), batch_inds=bi2[:Pg].numpy(), tmpps=tmpps.numpy(),
               grad_l_p=gl.numpy(), focals=focals0.numpy(), pps=pps0.numpy(), Rs=Rs0.numpy(), Ts=Ts0.numpy(),
               H=Hh, W=Ww, angthr=cam0.angThreshold(0.5))
    for case in ("fixedcam", "optcam"):
        opt = case == "optcam"
        cam_t = [t.clone().requires_grad_(opt) for t in (focals0, pps0, Rs0, Ts0)]
        cond_t = [t.clone().requires_grad_(True) for t in (poses, trans, dcond)]

        class FakeData:
            def get_grad_parameters(self, fids, device):
                return cond_t[0], cond_t[1], cond_t[2], None

            def get_camera_parameters(self, n, device):
                return cam_t[0], cam_t[1], cam_t[2], cam_t[3], Hh, Ww

        holder = types.SimpleNamespace(rasterizer=types.SimpleNamespace(cameras=cam0))
        sdf_full.zero_grad()
        comp.zero_grad()
        on = net_mod.OptimNetwork(sdf_full, comp, None, holder, None, conf=None)
        on.dataset = FakeData()
        on.info = {}
        on.TmpPs = tmpps.clone().requires_grad_(True)
        on.TmpPs.grad = gl.clone()
        camg = GoldCam(*cam_t, image_size=[(Ww, Hh)])
        pixh = torch.cat([col.view(-1, 1), row.view(-1, 1), torch.ones_like(col.view(-1, 1))], dim=-1).float()
        on.rays = camg.view_rays(pixh)
        on.col_inds, on.row_inds, on.batch_inds = col, row, bi2[:Pg]
        on.propagateTmpPsGrad(torch.arange(N), RATIO)
        out[case + "_invinfo"] = np.array(on.info['invInfo'])
        out[case + "_rays"] = on.rays.detach().numpy()
        named = [("sdf." + k, q) for k, q in sorted(sdf_full.named_parameters())] + \
                [("def." + k, q) for k, q in sorted(comp.named_parameters())] + \
                list(zip(("poses", "trans", "dcond"), cond_t))
        if opt:
            named += list(zip(("focals", "pps", "Rs", "Ts"), cam_t))
        for i, (k, q) in enumerate(named):
            if q.grad is None:
                continue
            out[case + "__" + k] = grad_digest(q.grad, i)
    np.savez(os.path.join(OUT, "propagate.npz"), **out)

    # ---------------------------------------------------------------- Seg3dLossless (reference)
    def query(points):
        q = points.reshape(-1, 3)
        r = q.norm(dim=1)
        val = r - 0.55 + 0.08 * torch.sin(7.0 * q[:, 0]) * torch.cos(5.0 * q[:, 1]) + 0.05 * q[:, 2]
        return val.reshape(1, 1, -1)

    ladder = [(9, 9, 9), (17, 17, 17), (33, 33, 33), (65, 65, 65)]
    eng = ref.seg3d.Seg3dLossless(query_func=query, b_min=[-1.0, -1.0, -1.0], b_max=[1.0, 1.0, 1.0],
                                  resolutions=ladder, align_corners=False, balance_value=0.0,
                                  visualize=False, debug=False, use_cuda_impl=False, faster=False)
    # record which voxels the reference queries by wrapping batch_eval
    queried = torch.zeros(65, 65, 65, dtype=torch.bool)
    orig = eng.batch_eval

    def spy(coords, **kw):
        c = coords[0]
        queried[c[:, 2], c[:, 1], c[:, 0]] = True
        return orig(coords, **kw)

    eng.batch_eval = spy
    grid = eng.forward()
    np.savez_compressed(os.path.join(OUT, "seg3d.npz"), **pack_grid(grid[0, 0], queried),
                        ladder=np.array(ladder),
                        spacing=np.array([eng.spacing_x, eng.spacing_y, eng.spacing_z]),
                        origin=np.array([eng.bx, eng.by, eng.bz]))
    # non-cubic ladder like the training one (15x21x9 -> 57x81x33)
    ladder2 = [(15, 21, 9), (29, 41, 17), (57, 81, 33)]
    eng2 = ref.seg3d.Seg3dLossless(query_func=query, b_min=[-0.9, -1.3, -0.5], b_max=[0.9, 0.9, 0.5],
                                   resolutions=ladder2, align_corners=False, balance_value=0.0,
                                   visualize=False, debug=False, use_cuda_impl=False, faster=False)
    queried2 = torch.zeros(33, 81, 57, dtype=torch.bool)
    orig2 = eng2.batch_eval

    def spy2(coords, **kw):
        c = coords[0]
        queried2[c[:, 2], c[:, 1], c[:, 0]] = True
        return orig2(coords, **kw)

    eng2.batch_eval = spy2
    grid2 = eng2.forward()
    np.savez_compressed(os.path.join(OUT, "seg3d_aniso.npz"), **pack_grid(grid2[0, 0], queried2),
                        ladder=np.array(ladder2))

    # ---------------------------------------------------------------- small helpers (a9, a18)
    from utils.FindSurfacePs import FindSurfacePs as ref_find
    g = torch.Generator().manual_seed(81)
    Nf, Hf, Wf, Kf, nF, nV = 2, 12, 10, 3, 40, 30
    p2f = torch.randint(-1, 2 * nF, (Nf, Hf, Wf, Kf), generator=g)        # packed face ids of 2 meshes, -1 = miss
    bary = torch.rand(Nf, Hf, Wf, Kf, 3, generator=g) - 0.15              # some non-positive weights
    TmpVs = torch.randn(nV, 3, generator=g)
    TmpFs = torch.randint(0, nV, (nF, 3), generator=g)
    frags = types.SimpleNamespace(pix_to_face=p2f, bary_coords=bary)
    bi_, ri_, ci_, ps_, finds_ = ref_find(TmpVs, TmpFs, frags)
    torch.manual_seed(82)
    pc = torch.randn(60, 3)
    sp = utils.sample_points(pc, 1.8, 0.01)
    xg = torch.linspace(0.0, 3.0, 25)
    qs = torch.randn(9, 4, generator=g)
    np.savez(os.path.join(OUT, "utils_misc.npz"), p2f=p2f.numpy(), bary=bary.numpy(), TmpVs=TmpVs.numpy(),
             TmpFs=TmpFs.numpy(), f_batch=bi_.numpy(), f_row=ri_.numpy(), f_col=ci_.numpy(), f_ps=ps_.numpy(),
             f_finds=finds_.numpy(), pc=pc.numpy(), sample=sp.numpy(), gm_x=xg.numpy(),
             gm=utils.GMRobustError(xg, 0.5).numpy(), gm_sq=utils.GMRobustError(xg, 0.5, True).numpy(),
             quat=qs.numpy(), quat_R=utils.quat2mat(qs).numpy(), apose0=utils.smpl_tmp_Apose(0),
             apose1=utils.smpl_tmp_Apose(1))

    # ---------------------------------------------------------------- batch_rodrigues
    g = torch.Generator().manual_seed(61)
    th = torch.randn(40, 3, generator=g)
    th[0] = 0.0
    np.savez(os.path.join(OUT, "rodrigues.npz"), theta=th.numpy(),
             R=ref.smpl_util.batch_rodrigues(th).numpy())
    print("golden vectors written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print("  %-20s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
