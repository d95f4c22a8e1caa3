"""One-process-per-GPU data parallelism for the hot path.

The reference has no distributed code at all; this is the B200-native design:

  * training  : frames (and their rays) are sharded across ranks; every rank holds full replicas
                of the three MLPs, the skin volume and the template mesh.  Gradients reach .grad
                from three backward calls per step (network.py:687, train.py:168, network.py:814)
                after one zero_grad, so ONE flat all-reduce (average) of {MLP params, per-frame
                leaf tensors} right after propagateTmpPsGrad covers them all (~15.5 MB over
                NVLink/NVSwitch: tens of microseconds; overlap is a non-goal).
  * rendering : frames are independent -> replicas only, no collective.
  * extraction: the final SDF grid is cut into x-slabs; each rank runs marching cubes on its slab
                plus a two-plane halo and the pieces are stitched into exactly the mesh a single
                GPU would produce (same canonical vertex / face order), with one all-gather of the
                counts and one of the payload.
Collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_frames(frame_ids, rank, world):
    """Contiguous, equal-sized shard of a step's frame ids (the global batch is N*world frames).
    Equal shards make 'mean over frames then mean over ranks' exact (network.py:617,637)."""
    n = len(frame_ids)
    if n % world != 0:
        raise ValueError("global batch of %d frames is not divisible by world size %d" % (n, world))
    per = n // world
    return frame_ids[rank * per:(rank + 1) * per]


class GradAllReduce:
    """Single bucketed all-reduce (average) of the gradients of `params` per step.

    The bucket is allocated once; parameters whose .grad is None contribute zeros (a frame-local
    leaf such as a latent code that this rank's frames did not touch)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.bucket = None

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1 or not self.params:
            return 0
        dev = self.params[0].device
        if self.bucket is None or self.bucket.device != dev:
            self.bucket = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.bucket[o:o + n].zero_()
            else:
                self.bucket[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)
        self.bucket.div_(dist.get_world_size(self.group))
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.bucket[o:o + n].view_as(p).clone()
            else:
                p.grad.copy_(self.bucket[o:o + n].view_as(p))
            o += n
        return self.numel * 4


def slab_range(nx, rank, world):
    """Cells [i0, i1) of the x axis owned by `rank` (nx grid planes -> nx-1 cells)."""
    cells = nx - 1
    base, rem = divmod(cells, world)
    i0 = rank * base + min(rank, rem)
    i1 = i0 + base + (1 if rank < rem else 0)
    return i0, i1


def _default_mc(sdf, step, origin, iso, i_offset):
    from . import ops
    return ops.marching_cubes(sdf, step[0], step[1], step[2], origin[0], origin[1], origin[2], iso,
                              i_offset)


def sharded_marching_cubes(sdf, step, origin, iso=0.0, rank=0, world=1, group=None, mc_fn=None):
    """Marching cubes over x-slabs, stitched to the single-device result.

    `sdf` [nx,ny,nz]: this rank only reads planes [i0, i1+2) of it (its cells plus the halo that
    makes the cells at plane i1 -- owned by the next rank -- valid), so callers may pass a tensor
    whose other planes were never evaluated.  Returns (vertices [V,3], faces [F,3] int64) of the
    WHOLE mesh on every rank, in the canonical order of the single-device kernel:
      - vertices owned by cells i < i1 of this rank keep their local order;
      - a reference to a vertex owned by a halo cell (plane i1) is rebased onto the next rank's
        first vertices: the halo cells produce the same vertices, in the same order, there.
    mc_fn(sdf, step, origin, iso, i_offset) -> (verts, faces) must be deterministic and canonical."""
    mc_fn = mc_fn or _default_mc
    nx = sdf.shape[0]
    i0, i1 = slab_range(nx, rank, world)
    last = (rank == world - 1)
    # the slab keeps the GLOBAL x index (i_offset) so positions are bit-identical to one device
    own_v, own_f = mc_fn(sdf[i0:i1 + 1].contiguous(), step, origin, iso, i0)      # cells i0 .. i1-1
    if last or world == 1:
        loc_v, loc_f = own_v, own_f
    else:
        loc_v, loc_f = mc_fn(sdf[i0:i1 + 2].contiguous(), step, origin, iso, i0)  # + halo cells at i1
    V_own, F_own = own_v.shape[0], own_f.shape[0]
    faces = loc_f[:F_own]
    verts = loc_v[:V_own]
    if world == 1:
        return verts, faces
    dev = sdf.device
    counts = torch.tensor([V_own, F_own], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = torch.stack(allc).cpu()
    voff = torch.cat([torch.zeros(1, dtype=torch.int64), allc[:, 0].cumsum(0)])
    # rebase: own vertices -> voff[rank] + id ; halo vertices -> voff[rank+1] + (id - V_own)
    f = faces.clone()
    halo = f >= V_own
    f = torch.where(f < 0, f, torch.where(halo, f - V_own + int(voff[min(rank + 1, world)]), f + int(voff[rank])))
    vmax, fmax = int(allc[:, 0].max()), int(allc[:, 1].max())
    vpad = torch.zeros((vmax, 3), dtype=verts.dtype, device=dev)
    vpad[:V_own] = verts
    fpad = torch.full((fmax, 3), -2, dtype=torch.int64, device=dev)
    fpad[:F_own] = f
    allv = [torch.empty_like(vpad) for _ in range(world)]
    allf = [torch.empty_like(fpad) for _ in range(world)]
    dist.all_gather(allv, vpad, group=group)
    dist.all_gather(allf, fpad, group=group)
    out_v = torch.cat([allv[r][:int(allc[r, 0])] for r in range(world)])
    out_f = torch.cat([allf[r][:int(allc[r, 1])] for r in range(world)])
    return out_v, out_f
