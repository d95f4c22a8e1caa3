// Pointwise stages of the tensor-core tracer (OptimizeSurfacePs, utils/FindSurfacePs.py:114-163).
// The dense layers run as tcgen05 split-BF16 GEMM launches (tc_gemm.cu); between them two small
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
  // borderline decisions: the tensor-core engine's f / D(p) carry up to ~2.4e-5 / ~1e-5 of error, so a
  // ray whose test could flip inside (eps_f, eps_a) is NOT decided here: it is appended to `recheck`
  // and re-tested by the fp32 FFMA engine (sr_trace_step_rev, test-only) before the update kernel runs.
  int* recheck;            // [P] or null (then the test below is final)
  int* recheck_count;
  float eps_f, eps_a;
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
      bool done = (fabsf(f) < a.tp.dthreshold) && (ang < a.tp.athreshold);
      if (a.recheck != nullptr) {
        const bool maybe = (fabsf(f) < a.tp.dthreshold + a.eps_f) && (ang < a.tp.athreshold + a.eps_a);
        const bool sure = (fabsf(f) < a.tp.dthreshold - a.eps_f) && (ang < a.tp.athreshold - a.eps_a);
        if (maybe && !sure) a.recheck[atomicAdd(a.recheck_count, 1)] = (int)gp;
        done = sure;
      }
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
  const unsigned char* converged;   // rays the fp32 re-test declared done after trace_mid (or null)
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
    if (a.converged != nullptr && a.converged[gp]) continue;
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

// ---- shading geometry from the tensor-core engine's outputs (infer path, network.py:356-361;
//      utils/utils.py:155-169): one warp per ray.
struct ShadePtArgs {
  long long P;
  const float* pts;
  const float* rays;
  const long long* batch_inds;
  const float* sdf4;   // [P*4][ld_s]: col 0 of rows 4p+1..4p+3 = grad f
  int ld_s;
  const float* off4;   // [P*4][3] translator offset (row 4p) and its tangents (rows 4p+1..3), or null
  sr_lbs_params lbs;
  int has_lbs;
  float* normals;
  float* crays;
  float* dpos;
  unsigned char* inv_ok;
};

__global__ void __launch_bounds__(256) shade_point_kernel(const __grid_constant__ ShadePtArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < a.P; i += nwarps) {
    float p[3], pp[3], off[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      p[j] = a.pts[i * 3 + j];
      if (a.off4) off[j] = a.off4[(i * 4) * 3 + j];
      pp[j] = __fadd_rn(p[j], off[j]);
    }
    float d[3], Mm[9];
    int ci[3];
    if (a.has_lbs) lbs_point(a.lbs, pp, a.batch_inds ? (int)a.batch_inds[i] : 0, d, Mm, ci);
    else {
#pragma unroll
      for (int j = 0; j < 3; ++j) d[j] = pp[j];
#pragma unroll
      for (int q = 0; q < 9; ++q) Mm[q] = (q % 4 == 0) ? 1.f : 0.f;
    }
    if (lane == 0) {
      // J = M (I + Joff), Joff[m][c] = d off_m / d p_c = off4[4i+1+c][m]
      float Q[9], m[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Q[3 * r + c] = (r == c ? 1.f : 0.f) + (a.off4 ? a.off4[(i * 4 + 1 + c) * 3 + r] : 0.f);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) m[3 * r + c] = Mm[3 * r] * Q[c] + Mm[3 * r + 1] * Q[3 + c] + Mm[3 * r + 2] * Q[6 + c];
      const float gx = a.sdf4[(i * 4 + 1) * a.ld_s], gy = a.sdf4[(i * 4 + 2) * a.ld_s], gz = a.sdf4[(i * 4 + 3) * a.ld_s];
      const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
      a.normals[i * 3] = gx / gn; a.normals[i * 3 + 1] = gy / gn; a.normals[i * 3 + 2] = gz / gn;
      const float c00 = m[4] * m[8] - m[5] * m[7], c01 = -m[3] * m[8] + m[5] * m[6], c02 = m[3] * m[7] - m[4] * m[6];
      const float c10 = -m[1] * m[8] + m[2] * m[7], c11 = m[0] * m[8] - m[2] * m[6], c12 = -m[0] * m[7] + m[1] * m[6];
      const float c20 = m[1] * m[5] - m[2] * m[4], c21 = -m[0] * m[5] + m[2] * m[3], c22 = m[0] * m[4] - m[1] * m[3];
      const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
      const bool ok = !(fabs((double)det) < 0.0001);
      const float vx = a.rays[i * 3], vy = a.rays[i * 3 + 1], vz = a.rays[i * 3 + 2];
      float rx = vx, ry = vy, rz = vz;
      if (ok) {
        rx = (c00 / det) * vx + (c10 / det) * vy + (c20 / det) * vz;
        ry = (c01 / det) * vx + (c11 / det) * vy + (c21 / det) * vz;
        rz = (c02 / det) * vx + (c12 / det) * vy + (c22 / det) * vz;
      }
      const float rn = sqrtf(rx * rx + ry * ry + rz * rz);
      a.crays[i * 3] = rx / rn; a.crays[i * 3 + 1] = ry / rn; a.crays[i * 3 + 2] = rz / rn;
      if (a.dpos) { a.dpos[i * 3] = d[0]; a.dpos[i * 3 + 1] = d[1]; a.dpos[i * 3 + 2] = d[2]; }
      if (a.inv_ok) a.inv_ok[i] = ok ? 1 : 0;
    }
  }
}

// rendering_input = cat([points, PE(view_dirs), normals, feature_vectors]) (RenderNet.py:73-74)
struct RenderEmbedArgs {
  long long P;
  const float* pts;
  const float* views;
  const float* normals;
  const float* feat;   // [P][feat_ld], first nfeat columns used (starting at feat_col0)
  int feat_ld, feat_col0, nfeat, feat_row_stride;
  int multires;
  float pw[16];
  float* out;
  int ld;
};
__global__ void render_embed_kernel(const __grid_constant__ RenderEmbedArgs a) {
  const long long total = a.P * (long long)a.ld;
  const int pe = 3 + 6 * a.multires;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % a.ld);
    const long long p = idx / a.ld;
    float v = 0.f;
    if (k < 3) v = a.pts[p * 3 + k];
    else if (k < 3 + pe) {
      const int q = k - 3;
      if (q < 3) v = a.views[p * 3 + q];
      else {
        const int b = (q - 3) / 6, w6 = (q - 3) % 6, j = w6 % 3;
        float sn, cs;
        sincosf(a.views[p * 3 + j] * (float)(1 << b), &sn, &cs);
        v = a.pw[b] * (w6 >= 3 ? cs : sn);
      }
    } else if (k < 3 + pe + 3) v = a.normals[p * 3 + (k - 3 - pe)];
    else if (k < 3 + pe + 3 + a.nfeat)
      v = a.feat[(size_t)p * a.feat_row_stride * a.feat_ld + a.feat_col0 + (k - 3 - pe - 3)];
    a.out[idx] = v;
  }
}

}  // namespace

extern "C" {

int sr_tc_shade_point(int64_t P, const float* pts, const float* rays, const int64_t* batch_inds,
                      const float* sdf4, int ld_s, const float* off4, const sr_lbs_params* lbs,
                      float* normals, float* crays, float* dpos, uint8_t* inv_ok, cudaStream_t s) {
  if (P <= 0 || !pts || !rays || !sdf4 || !normals || !crays) return SR_EINVAL;
  ShadePtArgs a;
  a.P = P; a.pts = pts; a.rays = rays; a.batch_inds = (const long long*)batch_inds; a.sdf4 = sdf4;
  a.ld_s = ld_s; a.off4 = off4; a.has_lbs = lbs ? 1 : 0;
  if (lbs) a.lbs = *lbs;
  a.normals = normals; a.crays = crays; a.dpos = dpos; a.inv_ok = inv_ok;
  shade_point_kernel<<<sr_grid_for(P * 32, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}

int sr_tc_render_embed(int64_t P, const float* pts, const float* views, const float* normals,
                       const float* feat, int feat_ld, int feat_col0, int nfeat, int feat_row_stride,
                       int multires, const float* pw, float* out, int ld, cudaStream_t s) {
  if (P <= 0 || !pts || !views || !normals || !out || !pw || (nfeat > 0 && !feat)) return SR_EINVAL;
  if (ld < 3 + 3 + 6 * multires + 3 + nfeat) return SR_EINVAL;
  RenderEmbedArgs a;
  a.P = P; a.pts = pts; a.views = views; a.normals = normals; a.feat = feat; a.feat_ld = feat_ld;
  a.feat_col0 = feat_col0; a.nfeat = nfeat; a.feat_row_stride = feat_row_stride; a.multires = multires;
  for (int i = 0; i < 16; ++i) a.pw[i] = i < multires ? pw[i] : 0.f;
  a.out = out; a.ld = ld;
  render_embed_kernel<<<sr_grid_for(P * (long long)ld, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}

int sr_tc_trace_mid(const int32_t* index, const int32_t* m_dev, int64_t P, const float* pts,
                    const float* rays, const int64_t* batch_inds, const float* f, const float* off,
                    const sr_lbs_params* lbs, const sr_trace_params* tp, int do_update,
                    uint8_t* converged, float* dsdf, float* ddef, int ld, float* aux,
                    int32_t* recheck, int32_t* recheck_count, float eps_f, float eps_a, cudaStream_t s) {
  if (!pts || !rays || !f || !tp || !converged || !dsdf || !aux || P <= 0 || ld < 8) return SR_EINVAL;
  if ((recheck != nullptr) != (recheck_count != nullptr) || eps_f < 0.f || eps_a < 0.f) return SR_EINVAL;
  MidArgs a;
  a.index = index; a.m_dev = m_dev; a.P = P; a.pts = pts; a.rays = rays;
  a.batch_inds = (const long long*)batch_inds; a.f = f; a.off = off;
  a.has_lbs = lbs ? 1 : 0;
  if (lbs) a.lbs = *lbs;
  a.tp = *tp; a.do_update = do_update; a.converged = converged; a.dsdf = dsdf; a.ddef = ddef; a.ld = ld;
  a.aux = aux; a.recheck = recheck; a.recheck_count = recheck_count; a.eps_f = eps_f; a.eps_a = eps_a;
  trace_mid_kernel<<<sr_grid_for(P * 32, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}

int sr_tc_trace_update(const int32_t* index, const int32_t* m_dev, int64_t P, float* pts,
                       const float* gs, int gs_ld, const float* gskip, int gk_ld, const float* gd,
                       int gd_ld, const float* aux, int mr_s, const float* pw_s, int mr_d,
                       const float* pw_d, int32_t* active_out, int32_t* counter_out,
                       const uint8_t* converged, cudaStream_t s) {
  if (!pts || !gs || !aux || !active_out || !counter_out || !pw_s || P <= 0) return SR_EINVAL;
  UpdArgs a;
  a.index = index; a.m_dev = m_dev; a.P = P; a.pts = pts; a.gs = gs; a.gs_ld = gs_ld; a.gskip = gskip;
  a.gk_ld = gk_ld; a.gd = gd; a.gd_ld = gd_ld; a.aux = aux; a.mr_s = mr_s; a.mr_d = mr_d;
  for (int i = 0; i < 16; ++i) { a.pw_s[i] = i < mr_s ? pw_s[i] : 0.f; a.pw_d[i] = (pw_d && i < mr_d) ? pw_d[i] : 0.f; }
  a.active_out = active_out; a.counter_out = counter_out; a.converged = converged;
  trace_update_kernel<<<sr_grid_for(P, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}
}
