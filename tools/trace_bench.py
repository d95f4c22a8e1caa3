"""CPU enqueue time vs device time of the tensor-core trace, eager and as a CUDA graph."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from selfreconcode_b200 import ops
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
sc = bench.build_scene(dev, frame_seed=0)
R = sc["rays"]
rays, init, bi = R["rays"].to(dev), R["init_pts"].to(dev), R["batch_inds"].to(dev)
out = {}
for mode in ("eager", "graph"):
    ops.GRAPHS_ENABLED = mode == "graph"
    ops._tc_trace_ctx.clear()
    for _ in range(3):
        bench.ray_part(sc, rays, init, bi)
    torch.cuda.synchronize()
    cpu, tot = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        p, c, rgb = bench.ray_part(sc, rays, init, bi)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        cpu.append(t1 - t0); tot.append(t2 - t0)
    out[mode] = {"cpu_enqueue_ms": 1e3 * sum(cpu) / 5, "total_ms": 1e3 * sum(tot) / 5,
                 "conv": int(c.sum()), "checksum": float(p.double().sum())}
print(json.dumps(out))
