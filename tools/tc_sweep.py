"""Per-launch fixed cost vs per-round cost of the CTA-pair layer kernel: M = r * 74 * 256 rows gives exactly r
rounds of 256x256 tiles x 2 column tiles on the 74 pairs; 20 back-to-back launches on rotating buffers."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfreconcode_b200 import ops
from selfreconcode_b200._lib import SR_ACT_SOFTPLUS100
dev = torch.device("cuda:0")
w = torch.randn(512, 512, device=dev) / 22.6
b = torch.zeros(512, device=dev)
W = ops.tc_pack_weights(w)
out = {}
for r in (1, 2, 3, 4, 6, 8):
    M = r * 37 * 256            # 37 row-pair-tiles x 2 column tiles = 74 pair tiles per round
    As = [ops.tc_pack_rows(torch.randn(M, 512, device=dev)) for _ in range(4)]
    for i in range(4):
        ops.tc_linear(As[i], W, b, M, 512, 512, 512, SR_ACT_SOFTPLUS100, K_next=512)
    torch.cuda.synchronize()
    keep = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        keep.append(ops.tc_linear(As[i & 3], W, b, M, 512, 512, 512, SR_ACT_SOFTPLUS100, K_next=512)[0])
        if len(keep) > 4:
            keep.pop(0)
    e1.record(); torch.cuda.synchronize()
    out["rounds%d_M%d_us" % (r, M)] = 1e3 * e0.elapsed_time(e1) / 20
    # the same 20 launches replayed as a CUDA graph (no CPU in the loop)
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for i in range(20):
            keep.append(ops.tc_linear(As[i & 3], W, b, M, 512, 512, 512, SR_ACT_SOFTPLUS100, K_next=512)[0])
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    out["rounds%d_graph_us" % r] = 1e3 * e0.elapsed_time(e1) / 20
    del g, keep
print(json.dumps(out))
