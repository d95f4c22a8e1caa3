"""Serialised (sync after every stage) time breakdown of the coarse-to-fine grid + MC part."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from selfreconcode_b200 import ops
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
sc = bench.build_scene(dev, frame_seed=0)
eng = bench.make_engine(sc, dev)
T = {}


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        T[name + "_n"] = T.get(name + "_n", 0) + 1
        return r
    return w


for _ in range(3):
    bench.mc_part(sc, eng)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    bench.mc_part(sc, eng)
torch.cuda.synchronize()
free = (time.perf_counter() - t0) / 5
import MCGpu
ops.interp2x3d_forward = timed("interp2x", ops.interp2x3d_forward)
ops.seg3d_candidates = timed("candidates", ops.seg3d_candidates)
qf = eng.query_func
eng.query_func = timed("query", qf)
eng.batch_eval = timed("batch_eval_total", eng.batch_eval)
MCGpu.mc_gpu = timed("mc_gpu", MCGpu.mc_gpu)
T.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    bench.mc_part(sc, eng)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / 5
out = {"free_running_ms": free * 1e3, "serialised_ms": tot * 1e3}
for k, v in T.items():
    out[k] = v / 5 * 1e3 if not k.endswith("_n") else v / 5
print(json.dumps(out))
