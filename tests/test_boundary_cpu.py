"""CPU: the reference-facing Python boundary.

Runs the import block of the reference's train.py / infer.py (the first-party names, train.py:3-11,
infer.py:3-11) against the drop-in directory, builds an OptimNetwork through `getOptNet` from a sequence on
disk, switches hierarchy level, and round-trips a `latest.pth`; the state_dict layout and the small helpers are
compared with fixtures generated from the unmodified reference (oracle/make_golden_r2.py boundary)."""
import os

import numpy as np
import pytest
import torch

from helpers import dropin, golden, rel_err


def _sequence(tmp, frames=8, H=24, W=20):
    dropin()
    from dataset import write_sequence
    from selfreconcode_b200 import synth
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(frames, H, W, 3, generator=g) * 2 - 1
    masks = (torch.rand(frames, H, W, generator=g) > 0.5).float()
    normals = torch.nn.functional.normalize(torch.randn(frames, H, W, 3, generator=g), dim=-1)
    poses, trans, _ = synth.make_frame_params(3, frames)
    cam = dict(fx=float(W), fy=float(W), cx=W / 2.0, cy=H / 2.0, quat=[0., 0., 0., 1.], T=[0., 0., 2.5])
    root = os.path.join(tmp, "seq")
    write_sequence(root, imgs.numpy(), masks.numpy(), poses.numpy(), trans.numpy(), np.zeros(10, np.float32), cam,
                   normals=normals.numpy())
    # cached LBS field: getOptNet then needs no SMPL model files (network.py:845-850)
    sk = synth.make_skinner(resolution=(9, 13, 5))
    torch.save({'ws': sk.ws, 'bmins': sk.b_min, 'bmaxs': sk.b_max, 'Js': sk.Js, 'parents': sk.parents,
                'init_pose': sk.init_pose, 'tmpBodyVs': torch.rand(30, 3, generator=g),
                'tmpBodyFs': torch.randint(0, 30, (40, 3), generator=g)}, os.path.join(root, 'initial_skinner_1.pth'))
    return root, imgs, masks, normals


def test_driver_import_block_resolves():
    dropin()
    ns = {}
    exec("from dataset.dataset import getDatasetAndLoader\n"
         "from model import getOptNet\n"
         "from MCAcc import Seg3dLossless\n"
         "import utils\n", ns)
    import model, utils, MCAcc  # noqa: E401
    for name in ("getTmpSdf", "OptimNetwork", "getOptNet", "initialLBSkinner", "RectifiedPerspectiveCameras",
                 "PointsRendererWithFrags"):                       # model/__init__.py:1-3
        assert hasattr(model, name), name
    for name in ("compute_lbswField", "FindSurfacePs", "OptimizeSurfacePs", "set_hierarchical_config", "save_model",
                 "load_model", "quat2mat", "annealing_weights", "GMRobustError", "smpl_tmp_Apose", "sample_points",
                 "compute_Jacobian", "compute_deformed_normals", "compute_cardinal_rays", "compute_netRender_color",
                 "FastDiff3x3MinvFunction", "DCTNullSpace", "DCTSpace", "compute_fnorms", "compute_vnorms",
                 "compute_face_areas"):                            # utils/__init__.py:1-3
        assert hasattr(utils, name), name
    for name in ("Seg3dLossless", "create_grid3D", "GridSamplerMine3dFunction"):   # MCAcc/__init__.py:1-3
        assert hasattr(MCAcc, name), name
    import model.network as mn
    assert mn.getOptNet is model.getOptNet and mn.OptimNetwork is model.OptimNetwork
    for meth in ("forward", "propagateTmpPsGrad", "initializeTmpSDF", "discretizeSDF", "infer", "computeTmpPcLoss",
                 "update_hierarchical_config"):
        assert callable(getattr(model.OptimNetwork, meth)), meth


def test_getoptnet_checkpoint_round_trip(tmp_path):
    root, imgs, masks, normals = _sequence(str(tmp_path))
    dropin()
    from dataset.dataset import getDatasetAndLoader
    from model import getOptNet
    import utils
    from selfreconcode_b200 import synth
    conf = synth.reference_config()
    dataset, loader = getDatasetAndLoader(root, {'deformer': 128, 'render': 16}, 2, True, 0, True, True, False)
    assert dataset.frame_num == 8 and (dataset.H, dataset.W) == (24, 20)
    idx, out = dataset[3]
    assert idx == 3 and out['img'].shape == (24, 20, 3) and out['mask'].shape == (24, 20)
    np.testing.assert_allclose(out['img'].numpy(), imgs[3].numpy(), atol=1.01 / 255)
    assert torch.equal(out['mask'], masks[3])
    np.testing.assert_allclose(out['normal'], normals[3].numpy(), atol=2.01 / 255)
    assert len(dataset.learnable_weights()) == 4      # two latent tables, poses, trans
    res = [(9, 13, 5), (17, 25, 9)]
    optNet, sdf_init = getOptNet(dataset, 2, None, None, res, torch.device("cpu"), conf)
    assert sdf_init == 1200 and optNet.remesh_intersect == 30 and hasattr(optNet, "dctnull")
    assert optNet.tmpBodyNs.shape == (30, 3)
    # state_dict layout = the reference's (keys and shapes; the LBS / engine buffers differ in size only)
    g = golden("boundary.npz")
    ref_keys = [str(k) for k in g["keys"]]
    sd = optNet.state_dict()
    assert list(sd.keys()) == ref_keys
    ref_shapes = dict(zip(ref_keys, [str(s) for s in g["shapes"]]))
    for k, v in sd.items():
        if k.startswith(("sdf.", "deformer.defs.0.", "netRender.")):
            assert str(tuple(v.shape)) == ref_shapes[k], k
    # hierarchy switch: new engine over the same box, queued confs (utils/utils.py:237-256)
    old = optNet.engine
    optNet, loader = utils.set_hierarchical_config(conf, 'medium', optNet, loader, [(9, 13, 5), (17, 25, 9), (33, 49, 17)])
    assert optNet.engine is not old and torch.equal(optNet.engine.b_min, old.b_min) and loader.batch_size == 2
    assert optNet.next_conf.get_float('color_weight') == 0.5
    assert optNet.next_train_conf.get_int('point_render.remesh_intersect') == 60
    # latest.pth round trip (utils/utils.py:257-316)
    path = os.path.join(str(tmp_path), "latest.pth")
    utils.save_model(path, 7, optNet, dataset)
    saved = torch.load(path, map_location="cpu", weights_only=False)
    assert set(saved) >= {"epoch", "model_state_dict", "poses", "trans", "shape", "dcond", "rcond", "focal_length",
                          "princeple_points", "cam2world_coord_quat", "world2cam_coord_trans"}
    before = {k: v.clone() for k, v in optNet.state_dict().items()}
    with torch.no_grad():
        for p in optNet.parameters():
            p.add_(1.0)
        dataset.poses.add_(1.0)
        ws0 = optNet.deformer.defs[1].ws.clone()
    optNet2, dataset2 = utils.load_model(path, optNet, dataset, torch.device("cpu"))
    for k, v in optNet2.state_dict().items():
        if 'engine.' in k or 'deformer.defs.1.ws' in k:
            continue
        assert torch.equal(v, before[k]), k
    assert torch.equal(optNet2.deformer.defs[1].ws, ws0), "the skin-weight volume is not restored from latest.pth"
    assert torch.equal(dataset2.poses, saved['poses']) and dataset2.poses.requires_grad
    assert dataset2.conds[0].requires_grad and dataset2.conds[1].requires_grad
    cams = optNet2.maskRender.rasterizer.cameras
    assert cams.R.shape == (2, 3, 3) and torch.allclose(cams.R[0], torch.diag(torch.tensor([-1., -1., 1.])))
    # sub-model substitution + prefix removal
    sdf_path = os.path.join(str(tmp_path), "sdf.pth")
    torch.save({k: v + 2.0 for k, v in optNet.sdf.state_dict().items()}, sdf_path)
    optNet3, _ = utils.load_model(path, optNet, dataset, torch.device("cpu"), subsdfmodel=sdf_path,
                                  model_rm_prefix=['netRender.'])
    k0 = next(iter(optNet.sdf.state_dict()))
    assert torch.equal(optNet3.state_dict()['sdf.' + k0], before['sdf.' + k0] + 2.0)


def test_batchframe_windows_and_samplers(tmp_path):
    root, *_ = _sequence(str(tmp_path), frames=12)
    dropin()
    from dataset.dataset import SceneDataset, RandomSampler, ShardedSampler
    ds = SceneDataset(root, {'deformer': 8})
    win, pos = ds.get_batchframe_data('poses', torch.tensor([0, 5, 11]), 5)
    assert win.shape == (3, 5, 24, 3) and pos.tolist() == [0, 2, 4]
    assert torch.equal(win[1], ds.poses[3:8]) and torch.equal(win[2], ds.poses[7:12])
    ds.video_segmented_index = [6]
    win, pos = ds.get_batchframe_data('trans', torch.tensor([5, 6]), 4)
    assert torch.equal(win[0], ds.trans[2:6]) and torch.equal(win[1], ds.trans[6:10]) and pos.tolist() == [3, 0]
    assert sorted(RandomSampler(ds, 1, True)) == list(range(12))
    shards = [list(ShardedSampler(ds, r, 8, shuffle=True, seed=3)) for r in range(8)]
    assert all(len(s) == 2 for s in shards)
    assert set(sum(shards, [])) == set(range(12))     # 16 draws cover the 12 frames, 4 repeated as padding
    p, t, c0, c1 = ds.get_grad_parameters(torch.tensor([1, 2]), "cpu")
    assert c0.shape == (2, 8) and c1 is None


def test_helpers_vs_reference_fixture():
    dropin()
    import utils
    from model.Deformer import compute_lbswField as field_model
    g = golden("boundary.npz")
    v, w = torch.from_numpy(g["lbsw_verts"]), torch.from_numpy(g["lbsw_ws"])
    box = ([-0.6, -0.7, -0.5], [0.6, 0.7, 0.5], (7, 9, 5))
    f = utils.compute_lbswField(*box, v, w, mean_neighbor=5, smooth_times=4)
    np.testing.assert_allclose(f.numpy(), g["lbsw_field"], atol=2e-6)
    np.testing.assert_allclose(field_model(*box, v, w, mean_neighbor=5, smooth_times=4).numpy(),
                               g["lbsw_field_model"], atol=2e-6)
    np.testing.assert_allclose(utils.DCTNullSpace(10, 30).numpy(), g["dctnull_10_30"], atol=1e-6)
    np.testing.assert_allclose(utils.DCTSpace(4, 20).numpy(), g["dctspace_4_20"], atol=1e-6)
    tv, tf = torch.from_numpy(g["tri_v"]), torch.from_numpy(g["tri_f"])
    np.testing.assert_allclose(utils.compute_face_areas(tv[None], tf).numpy(), g["face_areas"], atol=1e-6)
    np.testing.assert_allclose(utils.compute_fnorms(tv, tf).numpy(), g["fnorms"], atol=1e-5)
    from model.optim import vertex_face_pairs
    vid, fid = vertex_face_pairs(tf, 12)
    assert torch.equal(tf[fid].eq(vid.view(-1, 1)).any(1), torch.ones_like(vid, dtype=torch.bool))
    assert vid.numel() == tf.numel() and torch.equal(vid, vid.sort()[0])
