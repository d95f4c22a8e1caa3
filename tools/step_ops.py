"""Where the launches of one optimisation step come from: torch.profiler over one bench training step, kernel launches
attributed to the innermost frame inside this repository.  Writes gpurun_out/step_ops.txt."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    from selfreconcode_b200 import _lib
    _lib.load()
    sc = bench.build_scene(dev, 0)
    tr = bench.build_train(sc, dev, 0, 1)
    for _ in range(3):
        bench.train_step(tr)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        bench.train_step(tr)
        torch.cuda.synchronize()
    by_site = collections.Counter()
    by_op = collections.Counter()
    cpu_us = collections.Counter()
    for e in prof.events():
        n = len(e.kernels) if hasattr(e, "kernels") else 0
        if n == 0:
            continue
        site = "?"
        for fr in (e.stack or []):
            if ("selfreconcode_b200" in fr or "bench.py" in fr) and "site-packages" not in fr:
                site = fr.strip()
                break
        by_site[site] += n
        by_op[(site, e.name)] += n
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_ops.txt"), "w") as f:
        f.write("kernel launches of one step by source line (total %d)\n" % sum(by_site.values()))
        for s, n in by_site.most_common(80):
            f.write("%5d  %s\n" % (n, s.replace(ROOT + "/", "")))
            ops = [(k[1], v) for k, v in by_op.items() if k[0] == s]
            ops.sort(key=lambda t: -t[1])
            f.write("         " + ", ".join("%s x%d" % (a[:28], b) for a, b in ops[:6]) + "\n")
    print(open(os.path.join(ROOT, "gpurun_out", "step_ops.txt")).read()[:6000])


if __name__ == "__main__":
    main()
