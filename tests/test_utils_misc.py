"""Small torch-side helpers of the path (FindSurfacePs seed a9, sample_points / GMRobustError a18, quat2mat,
smpl_tmp_Apose) against values produced by the reference's own functions (oracle/make_golden.py)."""
import types

import numpy as np
import torch

import helpers as H


def test_find_surface_ps_matches_reference():
    H.dropin()
    import utils
    g = H.golden("utils_misc.npz")
    frags = types.SimpleNamespace(pix_to_face=torch.from_numpy(g["p2f"]), bary_coords=torch.from_numpy(g["bary"]))
    b, r, c, ps, finds = utils.FindSurfacePs(torch.from_numpy(g["TmpVs"]), torch.from_numpy(g["TmpFs"]), frags)
    assert np.array_equal(b.numpy(), g["f_batch"]) and np.array_equal(r.numpy(), g["f_row"])
    assert np.array_equal(c.numpy(), g["f_col"]) and np.array_equal(finds.numpy(), g["f_finds"])
    np.testing.assert_allclose(ps.numpy(), g["f_ps"], rtol=0, atol=1e-6)
    assert len(g["f_batch"]) > 20


def test_sampling_and_robust_error_match_reference():
    H.dropin()
    import utils
    g = H.golden("utils_misc.npz")
    torch.manual_seed(82)
    pc = torch.randn(60, 3)                                   # the golden script drew the cloud from the same stream
    assert np.array_equal(pc.numpy(), g["pc"])
    sp = utils.sample_points(pc, 1.8, 0.01)
    assert np.array_equal(sp.numpy(), g["sample"])          # same RNG stream, same op order
    x = torch.from_numpy(g["gm_x"])
    np.testing.assert_allclose(utils.GMRobustError(x, 0.5).numpy(), g["gm"], rtol=1e-6)
    np.testing.assert_allclose(utils.GMRobustError(x, 0.5, True).numpy(), g["gm_sq"], rtol=1e-6)
    np.testing.assert_allclose(utils.quat2mat(torch.from_numpy(g["quat"])).numpy(), g["quat_R"], atol=1e-6)
    assert np.array_equal(utils.smpl_tmp_Apose(0), g["apose0"]) and np.array_equal(utils.smpl_tmp_Apose(1), g["apose1"])
