"""One step of the bench workload with a reduced ray count, for ncu (never a bench number)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
sc = bench.build_scene(dev, 0)
eng = bench.make_engine(sc, dev)
R = sc["rays"]
n = R["rays"].shape[0]  # whole frame: the tensor-core engine needs realistic batch sizes
rays, init, bi = R["rays"][:n].to(dev), R["init_pts"][:n].to(dev), R["batch_inds"][:n].to(dev)
if "--mc-only" not in sys.argv:
    bench.ray_part(sc, rays, init, bi)
bench.mc_part(sc, eng)
torch.cuda.synchronize()
print("profile step done")
