"""TEST INFRASTRUCTURE -- numpy-facing ctypes wrapper of oracle/liboracle_c.so."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_c())
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def minv3x3(ms):
    ms = np.ascontiguousarray(ms, dtype=np.float32)
    n = ms.shape[0]
    inv = np.empty_like(ms)
    chk = np.empty(n, dtype=np.uint8)
    lib().orc_minv3x3_f32(_ptr(ms), _ptr(inv), _ptr(chk), C.c_int64(n))
    return inv, chk.astype(bool)


def minv3x3_bwd(grads, invs):
    g = np.ascontiguousarray(grads, dtype=np.float32)
    c = np.ascontiguousarray(invs, dtype=np.float32)
    out = np.empty_like(c)
    lib().orc_minv3x3_bwd_f32(_ptr(g), _ptr(c), _ptr(out), C.c_int64(c.shape[0]))
    return out


def marching_cubes(sdf, tri_table, iso=0.0, step=(1., 1., 1.), origin=(0., 0., 0.), i_offset=0):
    """sdf [nx,ny,nz] f32; tri_table int32 [256,16].  -> verts [V,3] f32, faces [F,3] i64 (canonical)."""
    sdf = np.ascontiguousarray(sdf, dtype=np.float32)
    tt = np.ascontiguousarray(tri_table, dtype=np.int32)
    nx, ny, nz = sdf.shape
    counts = np.zeros(2, dtype=np.int64)
    f = lib().orc_marching_cubes
    f.restype = C.c_int
    args = lambda v, fc: (_ptr(sdf), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_float(iso), _ptr(tt),
                          C.c_float(step[0]), C.c_float(step[1]), C.c_float(step[2]),
                          C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]), C.c_int(i_offset),
                          v, fc,
                          _ptr(counts))
    assert f(*args(C.c_void_p(0), C.c_void_p(0))) == 0
    verts = np.empty((int(counts[0]), 3), dtype=np.float32)
    faces = np.empty((int(counts[1]), 3), dtype=np.int64)
    assert f(*args(_ptr(verts), _ptr(faces))) == 0
    return verts, faces


def interp2x3d(inp, balance):
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    d, h, w = inp.shape
    out = np.empty((2 * d - 1, 2 * h - 1, 2 * w - 1), dtype=np.float32)
    bnd = np.empty(out.shape, dtype=np.uint8)
    lib().orc_interp2x3d_fwd(_ptr(inp), _ptr(out), _ptr(bnd), C.c_int(d), C.c_int(h), C.c_int(w),
                             C.c_float(balance))
    return out, bnd.astype(bool)


def grid_sample3d(inp, grid):
    """inp [C,D,H,W], grid [P,3] -> out [C,P], corner idx [P,3] int32."""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    Cc, D, H, W = inp.shape
    P = grid.shape[0]
    out = np.empty((Cc, P), dtype=np.float32)
    cidx = np.empty((P, 3), dtype=np.int32)
    lib().orc_grid_sample3d_fwd(_ptr(inp), _ptr(grid), _ptr(out), _ptr(cidx), C.c_int(Cc), C.c_int(D),
                                C.c_int(H), C.c_int(W), C.c_int64(P))
    return out, cidx
