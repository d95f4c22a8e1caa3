// Pointwise stages of the tensor-core tracer (OptimizeSurfacePs, utils/FindSurfacePs.py:114-163).
// The dense layers run as tcgen05 BF16x3 GEMM launches (tc_gemm.cu); between them two small
// kernels do everything that is per-ray:
//   trace_mid    after the forward sweeps: LBS of p + offset (warp per ray), convergence test,
//                loss, and the cotangents that seed the two backward sweeps
//   trace_update after the backward sweeps: chain rule through the positional encodings,
//                damped Newton step p <- p - loss/|g|^2 g, device-side append to the next list
#include "common.cuh"
#include "lbs.cuh"

namespace {

struct MidArgs {
  const int* index;        // active list (or null = identity)
  const int* m_dev;        // device-side count (or null -> P)
  long long P;
  const float* pts;        // [P,3] canonical points (global indexing)
  const float* rays;       // [P,3]
  const long long* batch_inds;
  const float* f;          // [M] sdf value of active row i
  const float* off;        // [M][3] translator offset (or null: identity deformer)
  sr_lbs_params lbs;
  int has_lbs;
  sr_trace_params tp;
  int do_update;
  unsigned char* converged;  // [P]
  float* dsdf;             // [M][ld] cotangent rows for the SDF backward sweep (col 0)
  float* ddef;             // [M][ld] cotangent rows for the translator backward sweep (cols 0..2)
  int ld;
  float* aux;              // [M][8]: loss, state, u[3]
};

__global__ void __launch_bounds__(256) trace_mid_kernel(const __grid_constant__ MidArgs a) {
  long long M = a.P;
  if (a.m_dev) { const long long md = *a.m_dev; M = md < M ? md : M; }
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < M; i += nwarps) {
    const long long gp = a.index ? (long long)a.index[i] : i;
    float p[3], pp[3], off[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      p[j] = a.pts[gp * 3 + j];
      if (a.off) off[j] = a.off[i * 3 + j];
      pp[j] = __fadd_rn(p[j], off[j]);
    }
    float d[3], Mm[9];
    int ci[3];
    if (a.has_lbs) {
      lbs_point(a.lbs, pp, a.batch_inds ? (int)a.batch_inds[gp] : 0, d, Mm, ci);
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) d[j] = pp[j];
#pragma unroll
      for (int q = 0; q < 9; ++q) Mm[q] = (q % 4 == 0) ? 1.f : 0.f;
    }
    if (lane == 0) {
      const float f = a.f[i];
      const float vx = a.rays[gp * 3], vy = a.rays[gp * 3 + 1], vz = a.rays[gp * 3 + 2];
      const float ux = d[0] - a.tp.cam_pos[0], uy = d[1] - a.tp.cam_pos[1], uz = d[2] - a.tp.cam_pos[2];
      const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
      const float n_up = sqrtf(cx * cx + cy * cy + cz * cz);
      const float n_u = sqrtf(ux * ux + uy * uy + uz * uz);
      const float sang = n_up / n_u;
      const float ang = asinf(sang) * 180.0f / 3.14159265358979323846f;
      const bool done = (fabsf(f) < a.tp.dthreshold) && (ang < a.tp.athreshold);
      float u[3] = {0.f, 0.f, 0.f}, loss = 0.f;
      int state = 0;
      if (done) {
        a.converged[gp] = 1;
      } else if (a.do_update) {
        state = 1;
        loss = a.tp.w1 * fabsf(f) + a.tp.w2 * fabsf(sang);
        float q[3] = {0.f, 0.f, 0.f};
        if (n_up > 0.f) {
          const float wx = vy * cz - vz * cy, wy = vz * cx - vx * cz, wz = vx * cy - vy * cx;
          const float i1 = 1.0f / (n_up * n_u), i2 = n_up / (n_u * n_u * n_u);
          q[0] = wx * i1 - ux * i2; q[1] = wy * i1 - uy * i2; q[2] = wz * i1 - uz * i2;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) u[j] = a.tp.w2 * (Mm[j] * q[0] + Mm[3 + j] * q[1] + Mm[6 + j] * q[2]);
      }
      float* ax = a.aux + i * 8;
      ax[0] = loss; ax[1] = __int_as_float(state); ax[2] = u[0]; ax[3] = u[1]; ax[4] = u[2];
    }
    // cotangent rows (the whole warp writes the ld-wide rows, zero padded)
    __syncwarp();  // lane 0's aux row is visible to the whole warp
    const float* ax = a.aux + i * 8;
    for (int k = lane; k < a.ld; k += 32) {
      float vs = 0.f, vd = 0.f;
      if (k == 0) {
        const float f = a.f[i];
        vs = (__float_as_int(ax[1]) == 1) ? a.tp.w1 * (f > 0.f ? 1.f : (f < 0.f ? -1.f : 0.f)) : 0.f;
      }
      if (k < 3) vd = ax[2 + k];
      a.dsdf[i * a.ld + k] = vs;
      if (a.ddef) a.ddef[i * a.ld + k] = vd;
    }
  }
}

struct UpdArgs {
  const int* index;
  const int* m_dev;
  long long P;
  float* pts;
  const float* gs;     // [M][gs_ld] dL/d(embedded sdf input), first 3+6L columns
  int gs_ld;
  const float* gskip;  // [M][gk_ld] skip-connection part (or null)
  int gk_ld;
  const float* gd;     // [M][gd_ld] dL/d(embedded translator input) (or null)
  int gd_ld;
  const float* aux;    // [M][8]
  int mr_s;
  float pw_s[16];
  int mr_d;
  float pw_d[16];
  int* active_out;
  int* counter_out;
};

__device__ __forceinline__ void pe_chain(const float* g, const float* gk, const float x[3], int multires,
                                         const float* pw, float out[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) out[j] += g[j] + (gk ? gk[j] : 0.f);
  float freq = 1.0f;
  for (int b = 0; b < multires; ++b, freq *= 2.0f) {
    const float w = pw[b] * freq;
    const int ks = 3 + 6 * b, kc = ks + 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float sn, cs;
      sincosf(x[j] * freq, &sn, &cs);
      const float gsn = g[ks + j] + (gk ? gk[ks + j] : 0.f), gcs = g[kc + j] + (gk ? gk[kc + j] : 0.f);
      out[j] += w * (cs * gsn - sn * gcs);
    }
  }
}

__global__ void __launch_bounds__(256) trace_update_kernel(const __grid_constant__ UpdArgs a) {
  long long M = a.P;
  if (a.m_dev) { const long long md = *a.m_dev; M = md < M ? md : M; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    const float* ax = a.aux + i * 8;
    if (__float_as_int(ax[1]) != 1) continue;
    const long long gp = a.index ? (long long)a.index[i] : i;
    const float x[3] = {a.pts[gp * 3], a.pts[gp * 3 + 1], a.pts[gp * 3 + 2]};
    float g[3] = {ax[2], ax[3], ax[4]};  // direct term u (d p' / d p = I + d off / d p)
    pe_chain(a.gs + i * a.gs_ld, a.gskip ? a.gskip + i * a.gk_ld : nullptr, x, a.mr_s, a.pw_s, g);
    if (a.gd) pe_chain(a.gd + i * a.gd_ld, nullptr, x, a.mr_d, a.pw_d, g);
    const float t = -ax[0] / (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) a.pts[gp * 3 + j] = x[j] + t * g[j];
    const int slot = atomicAdd(a.counter_out, 1);
    a.active_out[slot] = (int)gp;
  }
}

}  // namespace

extern "C" {

int sr_tc_trace_mid(const int32_t* index, const int32_t* m_dev, int64_t P, const float* pts,
                    const float* rays, const int64_t* batch_inds, const float* f, const float* off,
                    const sr_lbs_params* lbs, const sr_trace_params* tp, int do_update,
                    uint8_t* converged, float* dsdf, float* ddef, int ld, float* aux, cudaStream_t s) {
  if (!pts || !rays || !f || !tp || !converged || !dsdf || !aux || P <= 0 || ld < 8) return SR_EINVAL;
  MidArgs a;
  a.index = index; a.m_dev = m_dev; a.P = P; a.pts = pts; a.rays = rays;
  a.batch_inds = (const long long*)batch_inds; a.f = f; a.off = off;
  a.has_lbs = lbs ? 1 : 0;
  if (lbs) a.lbs = *lbs;
  a.tp = *tp; a.do_update = do_update; a.converged = converged; a.dsdf = dsdf; a.ddef = ddef; a.ld = ld;
  a.aux = aux;
  trace_mid_kernel<<<sr_grid_for(P * 32, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}

int sr_tc_trace_update(const int32_t* index, const int32_t* m_dev, int64_t P, float* pts,
                       const float* gs, int gs_ld, const float* gskip, int gk_ld, const float* gd,
                       int gd_ld, const float* aux, int mr_s, const float* pw_s, int mr_d,
                       const float* pw_d, int32_t* active_out, int32_t* counter_out, cudaStream_t s) {
  if (!pts || !gs || !aux || !active_out || !counter_out || !pw_s || P <= 0) return SR_EINVAL;
  UpdArgs a;
  a.index = index; a.m_dev = m_dev; a.P = P; a.pts = pts; a.gs = gs; a.gs_ld = gs_ld; a.gskip = gskip;
  a.gk_ld = gk_ld; a.gd = gd; a.gd_ld = gd_ld; a.aux = aux; a.mr_s = mr_s; a.mr_d = mr_d;
  for (int i = 0; i < 16; ++i) { a.pw_s[i] = i < mr_s ? pw_s[i] : 0.f; a.pw_d[i] = (pw_d && i < mr_d) ? pw_d[i] : 0.f; }
  a.active_out = active_out; a.counter_out = counter_out;
  trace_update_kernel<<<sr_grid_for(P, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}
}
