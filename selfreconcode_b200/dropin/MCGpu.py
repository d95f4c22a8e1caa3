"""Drop-in for the reference's pybind module `MCGpu` (MCGpu/MCGpu.cpp:14-60).

mc_gpu keeps the reference's positional signature and its legacy error convention (an EMPTY
LIST for a non-float32 grid, a device id outside [0,8) or non-positive dims, MCGpu.cpp:41-48;
RuntimeError for non-CUDA / non-contiguous input, the CHECK_INPUT macro at :30).  Unlike the
reference the output order is deterministic (canonical: vertices by owning edge key, faces by
voxel), buffers are sized exactly (the reference sizes them for 5% occupancy without an
overflow check, CudaKernels.cu:590-592) and no 12 B/voxel state volume is allocated.
"""
import torch

from selfreconcode_b200 import ops as _ops


def mc_init(device_id):
    """marching cube init a gpu device, not necessary for mc_gpu (kept for API parity)."""
    return None


def mc_gpu(sdfs, xstep=1.0, ystep=1.0, zstep=1.0, xmin=0.0, ymin=0.0, zmin=0.0, fTargetValue=0.0):
    """marching cube cuda global (CUDA): sdfs [NX,NY,NZ] -> [vertices [V,3] f32, faces [F,3] i64]."""
    if not sdfs.is_cuda:
        raise RuntimeError("sdfs must be a CUDA tensor")
    if not sdfs.is_contiguous():
        raise RuntimeError("sdfs must be contiguous")
    if sdfs.dtype != torch.float32:
        return []
    dev = sdfs.get_device()
    if dev < 0 or dev >= 8:
        return []
    if sdfs.dim() != 3 or min(sdfs.shape) <= 0:
        return []
    v, f = _ops.marching_cubes(sdfs.detach(), xstep, ystep, zstep, xmin, ymin, zmin, fTargetValue)
    return [v, f]
