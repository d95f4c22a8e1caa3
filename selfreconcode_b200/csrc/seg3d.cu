// Grid plumbing of the coarse-to-fine SDF evaluation (SURVEY.md row a20;
// reference: MCAcc/seg3d_lossless.py:266-314,348-372).
//
//   sr_seg3d_candidates: cand[z,y,x] = any(flag in the 3x3x3 neighbourhood, zero padded)
//                                      && !calculated[z*sz, y*sy, x*sx]
// replaces  smooth_conv3x3(is_boundary.float()) > 0  (a dense fp32 conv3d, :296) followed by
// is_boundary[coords_accum] = False (:299-301, `coords_accum` kept unique by a sort, :343-346):
// the set of already-evaluated lattice points is exactly the strided view of the final-grid
// `calculated` mask, so no coordinate list, no sort and no conv are needed.  The same kernel
// serves the conflict loop (flag = conflict mask, 27-neighbourhood, :354-372).
// Flags are sparse (a few % of the lattice), so the dilation is a SCATTER: the flag bytes are
// streamed 16 at a time (HBM-bound, 1 B per lattice point) and only set flags touch their 27
// neighbours; every writer stores the same value, so no atomics are needed.
//
//   sr_seg3d_gather / sr_seg3d_scatter: the per-pass glue around the query function
// (batch_eval's lattice -> world arithmetic :94-101 with the reference's rounding sequence, the
// gather of the interpolated values, `calculated[...] = True`, the write-back and the conflict
// test :330-346) as two kernels instead of ~25 torch launches.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
seg3d_dilate_kernel(const uint8_t* __restrict__ flag, const uint8_t* __restrict__ calculated,
                    uint8_t* __restrict__ cand, int D, int H, int W, int sz, int sy, int sx, int fH, int fW) {
  const long long total = (long long)D * H * W;
  const long long nvec = (total + 15) / 16;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (long long)gridDim.x * kThreads) {
    const long long base = v * 16;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (base + 16 <= total) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(flag + base));
      w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
      for (int b = 0; b < 16 && base + b < total; ++b) w[b >> 2] |= (uint32_t)(flag[base + b] != 0) << (8 * (b & 3));
    }
    if ((w[0] | w[1] | w[2] | w[3]) == 0u) continue;
    for (int b = 0; b < 16; ++b) {
      if (((w[b >> 2] >> (8 * (b & 3))) & 0xffu) == 0u) continue;
      const long long idx = base + b;
      const int x = (int)(idx % W), y = (int)((idx / W) % H), z = (int)(idx / ((long long)W * H));
      for (int dz = -1; dz <= 1; ++dz) {
        const int zz = z + dz;
        if (zz < 0 || zz >= D) continue;
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = y + dy;
          if (yy < 0 || yy >= H) continue;
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            if (calculated[((long long)zz * sz * fH + (long long)yy * sy) * fW + (long long)xx * sx] == 0)
              cand[((long long)zz * H + yy) * W + xx] = 1;
          }
        }
      }
    }
  }
}

struct GatherArgs {
  const long long* lin;
  long long n;
  int H, W, sz, sy, sx, fD, fH, fW;
  float bmin[3], bmax[3];
  const float* grid;
  float* points;
  float* interp;
  uint8_t* calculated;
};
__global__ void __launch_bounds__(kThreads) seg3d_gather_kernel(const __grid_constant__ GatherArgs a) {
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.n; i += (long long)gridDim.x * kThreads) {
    const long long l = a.lin[i];
    const int x = (int)(l % a.W), y = (int)((l / a.W) % a.H), z = (int)(l / ((long long)a.W * a.H));
    const int c[3] = {x * a.sx, y * a.sy, z * a.sz};       // coordinates on the final grid (x, y, z)
    const int res[3] = {a.fW, a.fH, a.fD};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // batch_eval (seg3d_lossless.py:99-101), one rounding per torch op:
      //   step = 1.0 / res; c2 = c / res + step / 2; p = c2 * (bmax - bmin) + bmin
      const float r = (float)res[d];
      const float step = __fdiv_rn(1.0f, r);
      const float c2 = __fadd_rn(__fdiv_rn((float)c[d], r), __fmul_rn(step, 0.5f));
      a.points[i * 3 + d] = __fadd_rn(__fmul_rn(c2, __fsub_rn(a.bmax[d], a.bmin[d])), a.bmin[d]);
    }
    a.interp[i] = a.grid[l];
    a.calculated[((long long)c[2] * a.fH + c[1]) * a.fW + c[0]] = 1;
  }
}

__global__ void __launch_bounds__(kThreads)
seg3d_scatter_kernel(const long long* __restrict__ lin, long long n, const float* __restrict__ values,
                     const float* __restrict__ interp, float balance, float* __restrict__ grid,
                     uint8_t* __restrict__ conflict, int* __restrict__ n_conflicts) {
  int mine = 0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const long long l = lin[i];
    const float t = values[i];
    grid[l] = t;
    // (interp - balance) * (true - balance) < 0  (seg3d_lossless.py:336)
    if (__fmul_rn(__fsub_rn(interp[i], balance), __fsub_rn(t, balance)) < 0.f) {
      conflict[l] = 1;
      ++mine;
    }
  }
  mine = sr_warp_sum_i(mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_conflicts, mine);
}
}  // namespace

extern "C" int sr_seg3d_candidates(const uint8_t* flag, const uint8_t* calculated, uint8_t* cand,
                                   int D, int H, int W, int sz, int sy, int sx, int fD, int fH,
                                   int fW, cudaStream_t s) {
  if (!flag || !calculated || !cand || D <= 0 || H <= 0 || W <= 0 || sz <= 0 || sy <= 0 || sx <= 0)
    return SR_EINVAL;
  if ((long long)(D - 1) * sz >= fD || (long long)(H - 1) * sy >= fH || (long long)(W - 1) * sx >= fW)
    return SR_EINVAL;
  if (((uintptr_t)flag & 15) != 0) return SR_EINVAL;
  const long long total = (long long)D * H * W;
  cudaError_t e = cudaMemsetAsync(cand, 0, (size_t)total, s);
  if (e != cudaSuccess) return (int)e;
  seg3d_dilate_kernel<<<sr_grid_for((total + 15) / 16, kThreads, 16), kThreads, 0, s>>>(
      flag, calculated, cand, D, H, W, sz, sy, sx, fH, fW);
  return sr_launch_status();
}

extern "C" int sr_seg3d_gather(const int64_t* lin, int64_t n, int H, int W, int sz, int sy, int sx, int fD,
                               int fH, int fW, const float* bmin, const float* bmax, const float* grid,
                               float* points, float* interp, uint8_t* calculated, cudaStream_t s) {
  if (!lin || n <= 0 || !bmin || !bmax || !grid || !points || !interp || !calculated || H <= 0 || W <= 0)
    return SR_EINVAL;
  GatherArgs a;
  a.lin = (const long long*)lin; a.n = n; a.H = H; a.W = W; a.sz = sz; a.sy = sy; a.sx = sx;
  a.fD = fD; a.fH = fH; a.fW = fW; a.grid = grid; a.points = points; a.interp = interp; a.calculated = calculated;
  for (int d = 0; d < 3; ++d) { a.bmin[d] = bmin[d]; a.bmax[d] = bmax[d]; }
  seg3d_gather_kernel<<<sr_grid_for(n, kThreads, 16), kThreads, 0, s>>>(a);
  return sr_launch_status();
}

extern "C" int sr_seg3d_scatter(const int64_t* lin, int64_t n, const float* values, const float* interp,
                                float balance, float* grid, uint8_t* conflict, int64_t conflict_bytes,
                                int32_t* n_conflicts, cudaStream_t s) {
  if (!lin || n <= 0 || !values || !interp || !grid || !conflict || !n_conflicts) return SR_EINVAL;
  cudaError_t e = cudaMemsetAsync(conflict, 0, (size_t)conflict_bytes, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(n_conflicts, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return (int)e;
  seg3d_scatter_kernel<<<sr_grid_for(n, kThreads, 16), kThreads, 0, s>>>((const long long*)lin, n, values, interp,
                                                                        balance, grid, conflict, n_conflicts);
  return sr_launch_status();
}
