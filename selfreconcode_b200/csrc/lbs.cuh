// Linear-blend skinning of one point by one warp (lane = bone), shared by the fused FFMA kernels
// (mlp_kernels.cu) and the tensor-core tracer's pointwise kernels (trace_tc.cu).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// LBS for one point, executed by one warp (lane = joint).   model/Deformer.py:205-233
// ---------------------------------------------------------------------------------------------
struct AxisF {
  int i0;
  float a[2];
  float mult;
  bool in0, in1;
};
__device__ __forceinline__ AxisF make_axis_f(float g, int size) {
  AxisF ax;
  const float prod = __fmul_rn(__fadd_rn(g, 1.0f), (float)size);
  float x = (float)(((double)prod - 1.0) / 2.0);
  const float hi = (float)(size - 1);
  if (!(x > 0.0f)) { ax.mult = 0.0f; x = 0.0f; }
  else if (x >= hi) { ax.mult = 0.0f; x = hi; }
  else ax.mult = 1.0f;
  const int i0 = (int)floorf(x);
  ax.i0 = i0;
  ax.a[0] = (float)(i0 + 1) - x;
  ax.a[1] = x - (float)i0;
  ax.in0 = i0 >= 0 && i0 < size;
  ax.in1 = (i0 + 1) >= 0 && (i0 + 1) < size;
  return ax;
}

// in : pp = p + offset, b;  out: d[3], M[9] = dD/dp' ; ci[3]
__device__ __forceinline__ void lbs_point(const sr_lbs_params& L, const float pp[3], int b,
                                          float d[3], float M[9], int ci[3]) {
  const int lane = threadIdx.x & 31;
  float nps[3], dn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float ext = __fsub_rn(L.bmax[j], L.bmin[j]);
    nps[j] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(pp[j], L.bmin[j])), ext), 1.0f);
    dn[j] = 2.0f / ext;
  }
  const AxisF ax = make_axis_f(nps[0], L.W), ay = make_axis_f(nps[1], L.H),
              az = make_axis_f(nps[2], L.D);
  ci[0] = ax.i0; ci[1] = ay.i0; ci[2] = az.i0;
  float wj = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  if (lane < 24) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int bx = k & 1, by = (k >> 1) & 1, bz = k >> 2;
      const bool ok = (bx ? ax.in1 : ax.in0) && (by ? ay.in1 : ay.in0) && (bz ? az.in1 : az.in0);
      if (ok) {
        const size_t vox = ((size_t)(az.i0 + bz) * L.H + (ay.i0 + by)) * L.W + (ax.i0 + bx);
        const float v = __ldg(L.ws_cl + vox * 24 + lane);
        const float w = ax.a[bx] * ay.a[by] * az.a[bz];
        wj = fmaf(v, w, wj);
        gx += v * (bx ? 1.f : -1.f) * ay.a[by] * az.a[bz];
        gy += v * (by ? 1.f : -1.f) * ax.a[bx] * az.a[bz];
        gz += v * (bz ? 1.f : -1.f) * ax.a[bx] * ay.a[by];
      }
    }
  }
  // d w_j / d p'  (grid-sampler coordinate gradient x d nps / d p')
  const float dwx = ax.mult * (gx * (float)L.W / 2.0f) * dn[0];
  const float dwy = ay.mult * (gy * (float)L.H / 2.0f) * dn[1];
  const float dwz = az.mult * (gz * (float)L.D / 2.0f) * dn[2];
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  if (lane < 24) {
    const float* A = L.A + ((size_t)b * 24 + lane) * 16;
    float h[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float a0 = __ldg(A + 4 * r), a1 = __ldg(A + 4 * r + 1), a2 = __ldg(A + 4 * r + 2),
                  a3 = __ldg(A + 4 * r + 3);
      h[r] = a0 * pp[0] + a1 * pp[1] + a2 * pp[2] + a3;
      acc[r] = wj * h[r];
      acc[3 + 3 * r + 0] = wj * a0 + h[r] * dwx;
      acc[3 + 3 * r + 1] = wj * a1 + h[r] * dwy;
      acc[3 + 3 * r + 2] = wj * a2 + h[r] * dwz;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = sr_warp_sum(acc[i]);
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = acc[r] + __ldg(L.trans + (size_t)b * 3 + r);
#pragma unroll
  for (int i = 0; i < 9; ++i) M[i] = acc[3 + i];
}

