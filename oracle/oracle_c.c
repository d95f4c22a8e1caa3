/*
 * oracle_c.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the integer / bit-level parts of the SelfRecon hot path, used
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.
 * The product (selfreconcode_b200/) never links, imports or calls anything in oracle/.
 *
 * Each function states the reference file:line it follows (jby1993/SelfReconCode @344b86f).
 * Parity pins: the reference has no golden vectors for this path; the
 * pins are (a) tests/golden/ *.npz produced by importing the reference's own Python on CPU
 * (oracle/make_golden.py), (b) the reference's CUDA kernels built into oracle/_ref/ and run
 * side by side on the GPU box (tests/test_ref_cuda_ab.py).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle_c.c -o liboracle_c.so -lm
 * (-ffp-contract=off: every fused multiply-add below is written explicitly with fmaf so
 * that the rounding matches what nvcc emits for the reference kernels.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Batched 3x3 inverse.  FastMinv/Matrix3x3InvKernels.cu:22-61 (forward), :64-104 (backward).
 * ------------------------------------------------------------------------------------------ */
void orc_minv3x3_f32(const float* ms, float* invs, uint8_t* checks, int64_t n) {
  for (int64_t q = 0; q < n; ++q) {
    const float* m = ms + 9 * q;
    float* o = invs + 9 * q;
    float c00 = m[4] * m[8] - m[5] * m[7];
    float c01 = -m[3] * m[8] + m[5] * m[6];
    float c02 = m[3] * m[7] - m[4] * m[6];
    float c10 = -m[1] * m[8] + m[2] * m[7];
    float c11 = m[0] * m[8] - m[2] * m[6];
    float c12 = -m[0] * m[7] + m[1] * m[6];
    float c20 = m[1] * m[5] - m[2] * m[4];
    float c21 = -m[0] * m[5] + m[2] * m[3];
    float c22 = m[0] * m[4] - m[1] * m[3];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (fabs((double)det) < 0.0001) {
      for (int i = 0; i < 9; ++i) o[i] = 0.f;
      checks[q] = 0;
    } else {
      o[0] = c00 / det; o[1] = c10 / det; o[2] = c20 / det;
      o[3] = c01 / det; o[4] = c11 / det; o[5] = c21 / det;
      o[6] = c02 / det; o[7] = c12 / det; o[8] = c22 / det;
      checks[q] = 1;
    }
  }
}

void orc_minv3x3_bwd_f32(const float* grads, const float* invs, float* outs, int64_t n) {
  for (int64_t q = 0; q < n; ++q) {
    const float *g = grads + 9 * q, *c = invs + 9 * q;
    float* o = outs + 9 * q;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        float acc = 0.f;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) acc += g[3 * i + j] * c[3 * i + a] * c[3 * b + j];
        o[3 * a + b] = -acc;
      }
  }
}

/* ------------------------------------------------------------------------------------------
 * Marching cubes with shared vertices.  MCGpu/CudaKernels.cu:304-521.
 *   Output in CANONICAL order (the reference's order is decided by atomicAdd races):
 *   vertices sorted by (i,j,k,dir) of the owning edge, faces by (voxel index, triangle#),
 *   winding reversed as d_conver_ijkd_to_pindex (:492-505) does, -1 for edges owned by a
 *   boundary-layer voxel (never created by d_mc_get_mesh_on_gpu, :451-459).
 * The triangulation is passed in by the caller (tests hold it as a fixture extracted from
 * the public-domain Bourke table) so that the oracle and the kernel do not share a table.
 * ------------------------------------------------------------------------------------------ */
static float orc_get_offset(float v1, float v2, float iso) { /* :304-313 */
  double delta = (double)(float)(v2 - v1);
  if (delta == 0.0) return 0.5f;
  return (float)((double)(float)(iso - v1) / delta);
}

static const int kConn[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6},
                                 {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int kOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                               {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
/* edge -> owner voxel offset + direction (the if/else ladder at :392-445) */
static const int kOwner[12][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {0, 1, 0, 0}, {0, 0, 0, 1},
                                  {0, 0, 1, 0}, {1, 0, 1, 1}, {0, 1, 1, 0}, {0, 0, 1, 1},
                                  {0, 0, 0, 2}, {1, 0, 0, 2}, {1, 1, 0, 2}, {0, 1, 0, 2}};

/* Returns 0 on success. counts[0]=V, counts[1]=F.  If verts/faces are NULL only counts. */
int orc_marching_cubes(const float* sdf, int nx, int ny, int nz, float iso, const int* tri_table,
                       float xs, float ys, float zs, float x0, float y0, float z0, int i_offset,
                       float* verts, int64_t* faces, int64_t* counts) {
  const int64_t nvox = (int64_t)nx * ny * nz;
  int32_t* vid = (int32_t*)malloc(sizeof(int32_t) * nvox * 3);
  if (!vid) return 1;
  memset(vid, 0xff, sizeof(int32_t) * nvox * 3);
#define SDF(i, j, k) sdf[((int64_t)(i) * ny + (j)) * nz + (k)]
  int64_t V = 0, F = 0;
  /* pass 1: vertices in (i,j,k,dir) order */
  for (int i = 0; i < nx - 1; ++i)
    for (int j = 0; j < ny - 1; ++j)
      for (int k = 0; k < nz - 1; ++k) {
        float v[8];
        int idx = 0;
        for (int c = 0; c < 8; ++c) {
          v[c] = SDF(i + kOff[c][0], j + kOff[c][1], k + kOff[c][2]);
          if (v[c] < iso) idx |= 1 << c;
        }
        if (idx == 0 || idx == 255) continue;
        const int own[3] = {0, 3, 8};
        for (int d = 0; d < 3; ++d) {
          const int e = own[d];
          const int a = kConn[e][0], b = kConn[e][1];
          if (((idx >> a) & 1) == ((idx >> b) & 1)) continue;
          const float t = orc_get_offset(v[a], v[b], iso);
          if (verts) {
            float px = (float)(i + i_offset), py = (float)j, pz = (float)k;
            if (e == 0) px = (float)(i + i_offset) + (0.0f + t);
            if (e == 3) py = (float)j + (1.0f - t); /* corner 3 -> corner 0, direction -y */
            if (e == 8) pz = (float)k + (0.0f + t);
            verts[3 * V + 0] = fmaf(px, xs, x0); /* d_scale_vertices: FMA-contracted */
            verts[3 * V + 1] = fmaf(py, ys, y0);
            verts[3 * V + 2] = fmaf(pz, zs, z0);
          }
          vid[(((int64_t)i * ny + j) * nz + k) * 3 + d] = (int32_t)V;
          ++V;
        }
      }
  /* pass 2: faces */
  for (int i = 0; i < nx - 1; ++i)
    for (int j = 0; j < ny - 1; ++j)
      for (int k = 0; k < nz - 1; ++k) {
        int idx = 0;
        for (int c = 0; c < 8; ++c)
          if (SDF(i + kOff[c][0], j + kOff[c][1], k + kOff[c][2]) < iso) idx |= 1 << c;
        const int* row = tri_table + idx * 16;
        for (int t = 0; t < 5; ++t) {
          if (row[3 * t] < 0) break;
          if (faces) {
            for (int c = 0; c < 3; ++c) {
              const int e = row[3 * t + c];
              const int oi = i + kOwner[e][0], oj = j + kOwner[e][1], ok = k + kOwner[e][2];
              faces[3 * F + (2 - c)] =
                  (int64_t)vid[(((int64_t)oi * ny + oj) * nz + ok) * 3 + kOwner[e][3]];
            }
          }
          ++F;
        }
      }
#undef SDF
  free(vid);
  counts[0] = V;
  counts[1] = F;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * 2x-1 upsample + boundary flag.  MCAcc/cuda/interp2x_boundary3d_kernel.cu:10-151.
 * ------------------------------------------------------------------------------------------ */
void orc_interp2x3d_fwd(const float* in, float* out, uint8_t* bnd, int d, int h, int w,
                        float balance) {
  const int od = 2 * d - 1, oh = 2 * h - 1, ow = 2 * w - 1;
#define IN(z, y, x) in[((int64_t)(z) * h + (y)) * w + (x)]
  for (int z = 0; z < od; ++z)
    for (int y = 0; y < oh; ++y)
      for (int x = 0; x < ow; ++x) {
        const int ex = x % 2 == 0, ey = y % 2 == 0, ez = z % 2 == 0;
        float v[8];
        int n = 0;
        if (ex && ey && ez) { v[n++] = IN(z / 2, y / 2, x / 2); }
        else if (ex && ez) { v[n++] = IN(z / 2, (y - 1) / 2, x / 2); v[n++] = IN(z / 2, (y + 1) / 2, x / 2); }
        else if (ey && ez) { v[n++] = IN(z / 2, y / 2, (x - 1) / 2); v[n++] = IN(z / 2, y / 2, (x + 1) / 2); }
        else if (ex && ey) { v[n++] = IN((z - 1) / 2, y / 2, x / 2); v[n++] = IN((z + 1) / 2, y / 2, x / 2); }
        else if (ez) {
          v[n++] = IN(z / 2, (y - 1) / 2, (x - 1) / 2); v[n++] = IN(z / 2, (y - 1) / 2, (x + 1) / 2);
          v[n++] = IN(z / 2, (y + 1) / 2, (x - 1) / 2); v[n++] = IN(z / 2, (y + 1) / 2, (x + 1) / 2);
        } else if (ex) {
          v[n++] = IN((z - 1) / 2, (y - 1) / 2, x / 2); v[n++] = IN((z + 1) / 2, (y - 1) / 2, x / 2);
          v[n++] = IN((z - 1) / 2, (y + 1) / 2, x / 2); v[n++] = IN((z + 1) / 2, (y + 1) / 2, x / 2);
        } else if (ey) {
          v[n++] = IN((z - 1) / 2, y / 2, (x - 1) / 2); v[n++] = IN((z + 1) / 2, y / 2, (x - 1) / 2);
          v[n++] = IN((z - 1) / 2, y / 2, (x + 1) / 2); v[n++] = IN((z + 1) / 2, y / 2, (x + 1) / 2);
        } else {
          for (int dz = -1; dz <= 1; dz += 2)
            for (int dy = -1; dy <= 1; dy += 2)
              for (int dx = -1; dx <= 1; dx += 2) v[n++] = IN((z + dz) / 2, (y + dy) / 2, (x + dx) / 2);
        }
        float s = v[0];
        int differ = 0;
        for (int t = 1; t < n; ++t) {
          s = s + v[t];
          if ((v[t] > balance) != (v[0] > balance)) differ = 1;
        }
        const int64_t o = ((int64_t)z * oh + y) * ow + x;
        out[o] = (float)((double)s / (double)n);
        bnd[o] = (uint8_t)differ;
      }
#undef IN
}

/* ------------------------------------------------------------------------------------------
 * Trilinear sample, border padding, align_corners=False: forward values and the corner
 * indices ("skinning indices").  MCAcc/cuda/GridSamplerMineKernel.cu:160-328.
 *   input [C][D][H][W] contiguous, grid [P][3] -> out [C][P], cidx [P][3]
 * ------------------------------------------------------------------------------------------ */
static float orc_unnorm_clip(float g, int size) {
  float prod = (g + 1.f) * (float)size;            /* float add, float mul (:210) */
  float x = (float)(((double)prod - 1.) / 2.);     /* double sub/div, rounded to float */
  float hi = (float)(size - 1);
  x = fmaxf(x, 0.f);                                /* clip_coordinates (:33-35) */
  x = fminf(hi, x);
  return x;
}

void orc_grid_sample3d_fwd(const float* input, const float* grid, float* out, int32_t* cidx, int C,
                           int D, int H, int W, int64_t P) {
  for (int64_t p = 0; p < P; ++p) {
    const float ix = orc_unnorm_clip(grid[3 * p], W), iy = orc_unnorm_clip(grid[3 * p + 1], H),
                iz = orc_unnorm_clip(grid[3 * p + 2], D);
    const int x0 = (int)floorf(ix), y0 = (int)floorf(iy), z0 = (int)floorf(iz);
    if (cidx) { cidx[3 * p] = x0; cidx[3 * p + 1] = y0; cidx[3 * p + 2] = z0; }
    const float ax[2] = {(float)(x0 + 1) - ix, ix - (float)x0};
    const float ay[2] = {(float)(y0 + 1) - iy, iy - (float)y0};
    const float az[2] = {(float)(z0 + 1) - iz, iz - (float)z0};
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int k = 0; k < 8; ++k) { /* tnw,tne,tsw,tse,bnw,bne,bsw,bse */
        const int bx = k & 1, by = (k >> 1) & 1, bz = k >> 2;
        const int x = x0 + bx, y = y0 + by, z = z0 + bz;
        if (x < 0 || x >= W || y < 0 || y >= H || z < 0 || z >= D) continue;
        const float w = ax[bx] * ay[by] * az[bz];
        acc = fmaf(input[(((int64_t)c * D + z) * H + y) * W + x], w, acc);
      }
      out[(int64_t)c * P + p] = acc;
    }
  }
}
