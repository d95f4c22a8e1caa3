"""CPU, world_size 2 over gloo: the N>1 host logic -- one flat gradient all-reduce per step and
the slab-sharded marching cubes stitch (marching cubes itself supplied by the C oracle here; on
GPUs it is the CUDA kernel)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, mc_tri_table


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grid(shape):
    ax = [np.linspace(-1, 1, s, dtype=np.float32) for s in shape]
    xx, yy, zz = np.meshgrid(*ax, indexing="ij")
    return (np.sqrt(xx * xx + yy * yy + zz * zz) - 0.7 + 0.06 * np.sin(5 * xx) * np.cos(3 * zz)).astype(np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from selfreconcode_b200 import parallel
    from oracle import c_api
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # ---- one flat all-reduce per step
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    latent = torch.nn.Parameter(torch.zeros(4, 6))       # per-frame leaf: only "own" rows get grads
    unused = torch.nn.Parameter(torch.zeros(3))          # .grad stays None on every rank
    frames = parallel.shard_frames(list(range(4)), rank, world)
    x = torch.arange(14, dtype=torch.float32).view(2, 7) * (rank + 1)
    loss = net(x).square().mean() + (latent[frames] * (rank + 1.0)).sum()
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    ar = parallel.GradAllReduce(list(net.parameters()) + [latent, unused])
    nbytes = ar()
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.numpy() for g in local])
    for i, p in enumerate(net.parameters()):
        mean = sum(gathered[k][i] for k in range(world)) / world
        np.testing.assert_allclose(p.grad.numpy(), mean, rtol=1e-6, atol=1e-7)
    lg = latent.grad.numpy()
    np.testing.assert_allclose(lg[:2], 0.5 * np.ones((2, 6)))   # frames of rank 0, weight 1, /2
    np.testing.assert_allclose(lg[2:], 1.0 * np.ones((2, 6)))   # frames of rank 1, weight 2, /2
    assert unused.grad is not None and float(unused.grad.abs().sum()) == 0.0
    assert nbytes == 4 * sum(p.numel() for p in list(net.parameters()) + [latent, unused])
    # ---- slab-sharded MC == single-device MC, bit for bit
    tt = mc_tri_table()

    def mc_fn(sdf, step, origin, iso, ioff):
        v, f = c_api.marching_cubes(sdf.numpy(), tt, iso, step, origin, ioff)
        return torch.from_numpy(v), torch.from_numpy(f)

    for shape in ((20, 17, 15), (33, 33, 33)):
        sdf = torch.from_numpy(_grid(shape))
        step, org = (0.11, 0.13, 0.09), (-1.0, -1.1, -0.7)
        v, f = parallel.sharded_marching_cubes(sdf, step, org, 0.0, rank, world, mc_fn=mc_fn)
        v1, f1 = mc_fn(sdf, step, org, 0.0, 0)
        assert torch.equal(f, f1), "stitched faces identical to the single-device canonical mesh"
        assert torch.equal(v, v1)
        # one emit pass + a count-only pass for the own / halo split; pieces gathered (unpadded) on rank 0 only
        def count_fn(s_, iso):
            a, b = mc_fn(s_, step, org, iso, 0)
            return a.shape[0], b.shape[0]
        v2, f2 = parallel.sharded_marching_cubes(sdf, step, org, 0.0, rank, world, mc_fn=mc_fn, count_fn=count_fn,
                                                 gather_to=0)
        if rank == 0:
            assert torch.equal(f2, f1) and torch.equal(v2, v1)
        else:
            assert v2.shape == (0, 3) and f2.shape == (0, 3)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, "ok"), (1, "ok")]
    assert all(p.exitcode == 0 for p in procs)


def test_slab_ranges_cover_all_cells():
    from selfreconcode_b200 import parallel
    for nx in (2, 9, 257, 513):
        for world in (1, 2, 3, 8):
            if world > nx - 1:
                continue
            cells = []
            for r in range(world):
                i0, i1 = parallel.slab_range(nx, r, world)
                cells += list(range(i0, i1))
            assert cells == list(range(nx - 1))
