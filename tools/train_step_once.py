"""One profiled optimisation step of the bench's training scene (for ncu):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/train_step_once.py
Two warm-up steps run outside the profiled range (cudaProfilerStart/Stop bracket exactly one step)."""
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
    for _ in range(2):
        bench.train_step(tr)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss, _ = bench.train_step(tr)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss %.6f rays %s" % (loss.item(), tr["net"].info["rayInfo"]))


if __name__ == "__main__":
    main()
