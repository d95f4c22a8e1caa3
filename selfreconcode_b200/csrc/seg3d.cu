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
// HBM-bound byte work: 1 B read (+26 L1/L2 hits) and 1 B written per lattice point.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
seg3d_candidates_kernel(const uint8_t* __restrict__ flag, const uint8_t* __restrict__ calculated,
                        uint8_t* __restrict__ cand, int D, int H, int W, int sz, int sy, int sx,
                        int fH, int fW) {
  const long long total = (long long)D * H * W;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int x = (int)(idx % W), y = (int)((idx / W) % H), z = (int)(idx / ((long long)W * H));
    bool any = false;
    for (int dz = -1; dz <= 1 && !any; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= D) continue;
      for (int dy = -1; dy <= 1 && !any; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const uint8_t* row = flag + ((long long)zz * H + yy) * W;
        any = (x > 0 && row[x - 1]) || row[x] || (x < W - 1 && row[x + 1]);
      }
    }
    if (any) any = calculated[((long long)z * sz * fH + (long long)y * sy) * fW + (long long)x * sx] == 0;
    cand[idx] = any ? 1 : 0;
  }
}
}  // namespace

extern "C" int sr_seg3d_candidates(const uint8_t* flag, const uint8_t* calculated, uint8_t* cand,
                                   int D, int H, int W, int sz, int sy, int sx, int fD, int fH,
                                   int fW, cudaStream_t s) {
  if (!flag || !calculated || !cand || D <= 0 || H <= 0 || W <= 0 || sz <= 0 || sy <= 0 || sx <= 0)
    return SR_EINVAL;
  if ((long long)(D - 1) * sz >= fD || (long long)(H - 1) * sy >= fH || (long long)(W - 1) * sx >= fW)
    return SR_EINVAL;
  seg3d_candidates_kernel<<<sr_grid_for((long long)D * H * W, kThreads, 16), kThreads, 0, s>>>(
      flag, calculated, cand, D, H, W, sz, sy, sx, fH, fW);
  return sr_launch_status();
}
