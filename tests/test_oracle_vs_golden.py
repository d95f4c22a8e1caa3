"""CPU: pins oracle/oracle.py against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  Sizes chosen so the whole CPU suite runs in a couple of minutes."""
import numpy as np
import torch

from helpers import (RATIO, SMPL_PARENTS, build_render, build_sdf_full, build_sdf_small,
                     build_skinner, build_translator, golden, plain_params, rel_err, sdf_params,
                     wn_params)
from oracle import oracle as O

TOL = 2e-6  # same algorithm, same fp32 torch kernels: only summation-order noise is expected


def test_embedder_and_annealing():
    g = golden("embedder.npz")
    x = torch.from_numpy(g["x"])
    np.testing.assert_array_equal(O.embed(x, 6).numpy(), g["embed6"])
    for r in (0.05, 0.3, 0.55, 1.0):
        ws = O.annealing_weights(6, r)
        np.testing.assert_array_equal(np.array(ws), g["aw_%g" % r])
        np.testing.assert_array_equal(O.embed(x, 6, ws).numpy(), g["embed6_r%g" % r])


def test_sdf_small_value_grad_feature():
    g = golden("sdf_small.npz")
    net = build_sdf_small(g)
    pts = torch.from_numpy(g["pts"])
    for r in (1.0, 0.4):
        s, gr, ft = O.sdf_value_and_grad(sdf_params(net), pts, 6, r, skip_in=(2,))
        assert rel_err(s.numpy(), g["sdf_r%g" % r].reshape(-1)) < TOL
        assert rel_err(gr.numpy(), g["grad_r%g" % r]) < 20 * TOL
        assert rel_err(ft.numpy(), g["feat_r%g" % r]) < TOL


def test_sdf_full_seeded():
    g = golden("sdf_full.npz")
    net = build_sdf_full(g)  # also proves the drop-in class initialises exactly like the reference
    s, gr, ft = O.sdf_value_and_grad(sdf_params(net), torch.from_numpy(g["pts"]), 6, 1.0)
    assert rel_err(s.numpy(), g["sdf"].reshape(-1)) < TOL
    assert rel_err(gr.numpy(), g["grad"]) < 20 * TOL
    assert rel_err(ft.numpy(), g["feat"]) < TOL


def _deform_setup(g):
    tr = build_translator(g)
    Js = torch.from_numpy(g["Js"])
    ipi = O.init_pose_inverse(torch.from_numpy(g["apose"]), Js, SMPL_PARENTS)
    np.testing.assert_allclose(ipi.numpy(), g["init_pose_inv"], atol=1e-6)
    poses, trans = torch.from_numpy(g["poses"]), torch.from_numpy(g["trans"])
    A, posed = O.bone_transforms(poses, Js, SMPL_PARENTS, ipi)
    lbs = dict(ws=torch.from_numpy(g["ws"]), bmin=torch.from_numpy(g["bmin"]),
               bmax=torch.from_numpy(g["bmax"]), A=A, trans=trans)
    return tr, lbs, posed


def test_deformer_composite_jacobian():
    g = golden("deform.npz")
    tr, lbs, posed = _deform_setup(g)
    assert rel_err(posed.numpy(), g["posed"]) < TOL
    pts, bi = torch.from_numpy(g["pts"]), torch.from_numpy(g["batch_inds"])
    dcond = torch.from_numpy(g["dcond"])
    r = float(g["def_ratio"])
    d_lbs = O.lbs_forward(lbs["ws"], lbs["bmin"], lbs["bmax"], lbs["A"], lbs["trans"], pts, bi)
    assert rel_err(d_lbs.numpy(), g["d_lbs"]) < TOL
    fn = lambda p: O.composite_deform(plain_params(tr), 6, r, dcond, lbs, p, bi)[0]
    d, J = O.jacobian(fn, pts)
    _, off = O.composite_deform(plain_params(tr), 6, r, dcond, lbs, pts, bi)
    assert rel_err(d.numpy(), g["d"]) < TOL
    assert rel_err(off.detach().numpy(), g["offset"]) < 10 * TOL
    assert rel_err(J.numpy(), g["jac"]) < 20 * TOL


def test_cardinal_rays():
    g, c = golden("deform.npz"), golden("cardinal.npz")
    tr, lbs, _ = _deform_setup(g)
    pts, bi = torch.from_numpy(g["pts"]), torch.from_numpy(g["batch_inds"])
    dcond = torch.from_numpy(g["dcond"])
    fn = lambda p: O.composite_deform(plain_params(tr), 6, float(g["def_ratio"]), dcond, lbs, p, bi)[0]
    cr, ds, J, ok = O.cardinal_rays(fn, pts, torch.from_numpy(c["rays"]))
    assert rel_err(cr.numpy(), c["crays"]) < 50 * TOL
    assert rel_err(ds.numpy(), c["ds"]) < TOL


def test_deformed_normals():
    g, n, gs = golden("deform.npz"), golden("normals.npz"), golden("sdf_full.npz")
    tr, lbs, _ = _deform_setup(g)
    sdf = build_sdf_full(gs)
    pts, bi = torch.from_numpy(g["pts"]), torch.from_numpy(g["batch_inds"])
    dcond = torch.from_numpy(g["dcond"])
    fn = lambda p: O.composite_deform(plain_params(tr), 6, float(g["def_ratio"]), dcond, lbs, p, bi)[0]
    sfn = lambda p: O.sdf_forward(sdf_params(sdf), p, 6, 1.0)[0]
    nn, ds = O.deformed_normals(sfn, fn, pts)
    assert rel_err(nn.numpy(), n["normals"]) < 50 * TOL
    assert rel_err(ds.numpy(), n["ds"]) < TOL


def test_render_net():
    g = golden("render.npz")
    rn = build_render(g)
    rgb = O.render_forward(wn_params(rn), torch.from_numpy(g["pts"]), torch.from_numpy(g["normals"]),
                           torch.from_numpy(g["views"]), torch.from_numpy(g["feat"]), 4, 1.0)
    assert rel_err(rgb.numpy(), g["rgb"]) < TOL


def test_rodrigues():
    g = golden("rodrigues.npz")
    R = O.batch_rodrigues(torch.from_numpy(g["theta"]))
    np.testing.assert_allclose(R.numpy(), g["R"], atol=1e-6)


def test_optimize_surface_ps():
    t, g, gs = golden("trace.npz"), golden("deform.npz"), golden("sdf_full.npz")
    sdf = build_sdf_full(gs)
    tr, lbs, _ = _deform_setup(g)
    dcond = torch.from_numpy(g["dcond"])
    sp = sdf_params(sdf)
    sdf_fn = lambda p: O.sdf_forward(sp, p, 6, 1.0)[0]
    def_fn = lambda p, b: O.composite_deform(plain_params(tr), 6, float(g["def_ratio"]), dcond, lbs, p, b)[0]
    rays, start = torch.from_numpy(t["rays"]), torch.from_numpy(t["start"])
    bi, cam = torch.from_numpy(t["batch_inds"]), torch.from_numpy(t["cam_pos"])
    for name, (dth, times) in {"train": (5e-5, 10), "infer": (1e-4, 30)}.items():
        p, conv, _ = O.optimize_surface_ps(cam, rays, start, bi, sdf_fn, def_fn, dth,
                                           float(t["athreshold"]), 3.05, 1.0, times)
        ref_conv = t["conv_" + name]
        # convergence is a threshold on values with ~1e-6 noise: allow a borderline ray or two
        assert (conv.numpy() != ref_conv).sum() <= 2
        # the iteration is a contraction: every ray (converged or not) must land on the same point
        assert np.abs(p.numpy() - t["pts_" + name]).max() < 2e-5


def _unpack(g):
    shape = tuple(int(v) for v in g["shape"])
    n = int(np.prod(shape))
    q = np.unpackbits(g["queried"])[:n].astype(bool).reshape(shape)
    sign = np.unpackbits(g["sign"])[:n].astype(bool).reshape(shape)
    return shape, q, sign


def _query(points):
    q = points.reshape(-1, 3)
    r = q.norm(dim=1)
    return r - 0.55 + 0.08 * torch.sin(7.0 * q[:, 0]) * torch.cos(5.0 * q[:, 1]) + 0.05 * q[:, 2]


def test_seg3d_lossless_cubic_and_anisotropic():
    for name, bmin, bmax in (("seg3d.npz", [-1.0] * 3, [1.0] * 3),
                             ("seg3d_aniso.npz", [-0.9, -1.3, -0.5], [0.9, 0.9, 0.5])):
        g = golden(name)
        shape, q_ref, sign_ref = _unpack(g)
        ladder = [tuple(int(v) for v in r) for r in g["ladder"]]
        grid, calc = O.seg3d_forward(_query, bmin, bmax, ladder, 0.0)
        assert tuple(grid.shape) == shape
        np.testing.assert_array_equal(calc.numpy(), q_ref)              # same set of queried voxels
        # values at queried voxels: identical up to the 1-ulp vectorised-vs-tail difference of
        # torch's CPU sin/cos when the query batch is composed in a different order
        np.testing.assert_allclose(grid.numpy()[q_ref], g["values_at_queried"], atol=5e-7)
        np.testing.assert_array_equal(grid.numpy() > 0.0, sign_ref)     # what MC consumes
        np.testing.assert_allclose(grid.numpy().reshape(-1)[g["interp_idx"]], g["interp_val"],
                                   atol=2e-6)
