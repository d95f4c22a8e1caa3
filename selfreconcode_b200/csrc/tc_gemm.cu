// Tensor-core engine for the dense layers: split-BF16 GEMM on tcgen05 with fp32 accumulation in TMEM
// (SURVEY.md section 7 "hard parts": the ray finder tests |f| < 5e-5; a single BF16 or TF32 pass does
// not hold that).  Every fp32 operand is split into bf16 planes x = b1 + b2 (+ b3) and the product is
// assembled from the leading cross terms:
//     2 planes / 3 MMAs (default):  x*w ~= b1w1 + b1w2 + b2w1         (dropped: <= 2^-16 relative)
//     3 planes / 6 MMAs (-DSR_TC_PLANES=3): + b2w2 + b1w3 + b3w1      (dropped: <= 2^-24 relative)
// Measured on B200 against an fp64 evaluation of the 8x512 SDF (tools/tc_terms.py): the tensor core
// adds each MMA into the fp32 accumulator with truncation, so the error of a K=512 layer is dominated
// by the NUMBER of accumulations (K/16 per term), not by the dropped terms -- 3 terms: max |err|
// 2.4e-5, 6 terms: 3.4e-5 (FFMA engine: 2.3e-6; parity bar 1e-4).  Fewer terms are both faster
// and closer, hence the default.
//
// One launch = one layer  C[M x N] = act(A[M x K] * W^T + b)  over all row tiles:
//   * operands live in global memory already in the canonical (no-swizzle, K-major) shared
//     memory layout of the UMMA descriptors -- 128-byte core matrices (8 rows x 8 bf16), tiles
//     of 128 (or 256) rows x 32 k, the split planes of a tile contiguous -- so ONE TMA bulk
//     copy (cp.async.bulk + mbarrier) per operand per stage lands a ready-to-use tile and no
//     tensor map is needed.  The previous layer's epilogue writes its output directly in that
//     layout (activations stay L2-resident between layers for the batch sizes used here);
//   * warp-specialised, persistent CTAs: warp 0 = TMA producer, warp 1 = single-thread
//     tcgen05.mma issuer (+ TMEM alloc), warps 2-9 = epilogue (tcgen05.ld -> bias, activation,
//     forward-mode tangent scaling or reverse-mode act' multiply, re-split to bf16 planes, tiled
//     store); the epilogue is specialised at compile time per (activation, rows-per-point, mode);
//   * 4-stage shared-memory ring (48 KB per stage), two 256-column fp32 accumulators in TMEM so
//     the epilogue of tile i overlaps the MMAs of tile i+1.
// The FFMA engine (mlp_kernels.cu) stays the accuracy reference; tests compare both.
#include <cstdlib>

#include "tc_common.cuh"

namespace sr_tc {

struct LayerArgs {
  const __nv_bfloat16* A;   // tiled activations  [MT][KC][3][128x32]
  const __nv_bfloat16* W;   // tiled weights      [NT][KC][planes][256x32]
  const __nv_bfloat16* Wp;  // the same weights in the CTA-pair layout [NT][KC][half][planes][128x32]
  const float* bias;        // [NT*256]
  long long M;              // valid rows
  int MT, NT, KC;           // row tiles, col tiles, k chunks (K = 32*KC)
  int n_gemm;               // columns the GEMM produces (<= NT*256); the MMA N of the last tile shrinks to it
  int n;                    // valid output columns
  int ch;                   // rows per point: 1 (value only) or 4 (value + 3 tangents)
  // outputs (either may be null)
  __nv_bfloat16* A_next;    // tiled, KCn chunks: activations for the next layer
  int KCn;
  float scale;              // 1, or 1/sqrt(2) when the next layer is the skip layer
  const float* skip_src;    // fp32 [M][skip_ld] embedded input appended after column n (or null)
  int skip_n, skip_ld;
  float* out;               // fp32 row-major [M][out_ld] (last layer), columns [0, n)
  int out_ld;
  float* dstash;            // fp32 row-major [M][NT*256] act'(z) of value rows (reverse mode) or null
  // reverse sweep (MUL kernels): out = acc * act'(z_prev), act' recomputed from the previous layer's
  // stored output a (the tiles the forward pass wrote for this layer's input, scaled by 1/mul_inv_scale):
  // softplus100: 1 - exp(-100 a), relu: a > 0
  const __nv_bfloat16* mul_tiles;
  int mul_KC;
  float mul_inv_scale;
  int out_col0;             // `out` receives columns [out_col0, out_col0 + out_n) of the result
  int out_n;
  const int* m_dev;         // optional device-side row count (active rays); M is the upper bound
};

template <int ACT>
__device__ __forceinline__ float act_fn(float z, float& d) {
  if constexpr (ACT == SR_ACT_SOFTPLUS100) {
    const float bz = z * 100.0f;
    const float e = __expf(fminf(bz, 20.0f));
    const float sp = __logf(1.0f + e) * 0.01f;
    const float dd = __fdividef(e, e + 1.0f);
    d = bz > 20.0f ? 1.0f : dd;
    return bz > 20.0f ? z : sp;
  } else if constexpr (ACT == SR_ACT_RELU) {
    d = z > 0.f ? 1.f : 0.f;
    return fmaxf(z, 0.f);
  } else if constexpr (ACT == SR_ACT_TANH) {
    const float t = tanhf(z);
    d = 1.f - t * t;
    return t;
  } else {
    d = 1.f;
    return z;
  }
}

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// value only (the common epilogue: no tangent rows, no stash): 2 MUFU per softplus
template <int ACT>
__device__ __forceinline__ float act_val(float z) {
  if constexpr (ACT == SR_ACT_SOFTPLUS100) {
    // softplus(beta = 100, threshold 20): log(1 + exp(100 z)) / 100, z itself above the threshold
    const float t = z * 144.26950408889634f;                       // 100 z log2(e)
    const float e = fast_ex2(fminf(t, 28.853900817779268f));       // exp(min(100 z, 20))
    const float sp = fast_lg2(1.0f + e) * 0.006931471805599453f;   // ln(2) / 100
    return t > 28.853900817779268f ? z : sp;
  } else if constexpr (ACT == SR_ACT_RELU) {
    return fmaxf(z, 0.f);
  } else if constexpr (ACT == SR_ACT_TANH) {
    return tanhf(z);
  } else {
    return z;
  }
}
// act'(z) recovered from the activation value a = act(z)
template <int ACT>
__device__ __forceinline__ float dact_from_val(float a) {
  if constexpr (ACT == SR_ACT_SOFTPLUS100) return 1.0f - fast_ex2(a * -144.26950408889634f);
  else if constexpr (ACT == SR_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  else if constexpr (ACT == SR_ACT_TANH) return 1.f - a * a;
  else return 1.f;
}

// two fp32 values -> three packed bf16 pairs (element 0 in the low half = lower address)
__device__ __forceinline__ void split3x2(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  p1 = *reinterpret_cast<uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(p1 << 16), r1 = x1 - __uint_as_float(p1 & 0xffff0000u);
  h = __floats2bfloat162_rn(r0, r1);
  p2 = *reinterpret_cast<uint32_t*>(&h);
  if constexpr (kPlanes == 3) {
    const float s0 = r0 - __uint_as_float(p2 << 16), s1 = r1 - __uint_as_float(p2 & 0xffff0000u);
    h = __floats2bfloat162_rn(s0, s1);
    p3 = *reinterpret_cast<uint32_t*>(&h);
  } else {
    p3 = 0u;
  }
}

struct EpiRow {
  long long mt, row;
  int nt, row_in_tile, lane;
  bool row_ok, is_val;
  size_t ds_ld;
};

// One 32-column chunk of the accumulator (this thread: one row): bias / activation / tangent
// scaling (forward) or act' multiply (reverse), then the fp32 outputs, the act' stash and the
// re-split bf16 tile of the next layer.  `live` = the chunk holds GEMM columns (else only the
// zero padding / skip-connection columns of the next layer's input are produced).
template <int ACT, int CH, bool MUL, bool PF = true>
__device__ __forceinline__ void epi_chunk(const LayerArgs& a, const EpiRow& r, uint32_t (&v)[32], int chunk,
                                          bool live, bool wait_v = false) {
  const int c0 = r.nt * BN + chunk * 32;
  float o[32];
  if (live) {
    if constexpr (MUL) {
      // reverse sweep: delta * act'(z) for the columns that continue; columns >= n (the skip part of
      // a skip layer's input gradient) pass through unscaled.  ACT is the PREVIOUS layer's activation.
      const __nv_bfloat16* mt = a.mul_tiles + a_tile_off(r.mt, c0 >> 5, a.mul_KC, 0) +
                                (size_t)(r.row_in_tile >> 3) * 64 + (r.row_in_tile & 7) * 8;
      const float kk = -144.26950408889634f * a.mul_inv_scale;   // -100 log2(e) / scale
      // optional fp32 act'(z) of the previous layer's VALUE rows (written by its forward launch as `dstash`)
      // rows past M read row 0's entries (discarded below): the predicate stays warp-uniform, so the warp is
      // converged at the .aligned TMEM wait that follows the loads
      const long long srow = r.row_ok ? (CH == 4 ? (r.row & ~3LL) : r.row) : 0;
      const float* stash = a.dstash == nullptr ? nullptr : a.dstash + (size_t)srow * r.ds_ld + c0;
      const bool use_stash = (ACT == SR_ACT_SOFTPLUS100) && stash != nullptr;
      // all global operands of the chunk first (8 x 16 B of activation tiles, 8 x 16 B of stash): 16 loads in flight
      // per thread -- with two epilogue warps per scheduler nothing else hides their latency
      // (PF = false: the 16-warp build of the reverse kernels has 96 registers -- operands are requested per group of
      //  8 columns and the extra warps hide the latency instead)
      uint4 q0[PF ? 4 : 1], q1[PF ? 4 : 1];
      float st[PF ? 32 : 1];
      if constexpr (PF) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          q0[g] = __ldg(reinterpret_cast<const uint4*>(mt + (size_t)g * (BM * 8)));
          q1[g] = __ldg(reinterpret_cast<const uint4*>(mt + (size_t)g * (BM * 8) + A_PLANE));
        }
        if (use_stash) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(stash) + j4);
            st[4 * j4] = t.x; st[4 * j4 + 1] = t.y; st[4 * j4 + 2] = t.z; st[4 * j4 + 3] = t.w;
          }
        }
      }
      __syncwarp();
      if (wait_v) tmem_wait(v);     // the accumulator chunk was requested by the caller before this function
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 a0, a1;
        float sg[8];
        if constexpr (PF) {
          a0 = q0[g]; a1 = q1[g];
        } else {
          a0 = __ldg(reinterpret_cast<const uint4*>(mt + (size_t)g * (BM * 8)));
          a1 = __ldg(reinterpret_cast<const uint4*>(mt + (size_t)g * (BM * 8) + A_PLANE));
          if (use_stash) {
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(stash) + 2 * g);
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(stash) + 2 * g + 1);
            sg[0] = t0.x; sg[1] = t0.y; sg[2] = t0.z; sg[3] = t0.w; sg[4] = t1.x; sg[5] = t1.y; sg[6] = t1.z; sg[7] = t1.w;
          }
        }
        const uint32_t w0[4] = {a0.x, a0.y, a0.z, a0.w}, w1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = g * 8 + e;
          const uint32_t h0 = w0[e >> 1], h1 = w1[e >> 1];
          const float as = (e & 1) ? __uint_as_float(h0 & 0xffff0000u) + __uint_as_float(h1 & 0xffff0000u)
                                   : __uint_as_float(h0 << 16) + __uint_as_float(h1 << 16);
          float val = __uint_as_float(v[j]);
          if constexpr (CH == 1) {
            float d;
            if constexpr (ACT == SR_ACT_SOFTPLUS100) {
              // training: act'(z) kept in fp32 by the forward sweep (recomputing it from the 16-bit-mantissa
              // activation tiles costs 100 x 2^-17 relative on 1 - act'); the tracer recomputes (no stash traffic)
              d = use_stash ? (PF ? st[PF ? j : 0] : sg[e]) : 1.0f - fast_ex2(kk * as);
            } else if constexpr (ACT == SR_ACT_RELU) d = as > 0.f ? 1.f : 0.f;
            else d = 1.f;
            if (c0 + j < a.n) val = r.row_ok ? val * d : 0.f;
          } else {
            // Reverse sweep over forward-mode rows (value + 3 tangents per point; training: the cotangents of
            // f AND of grad f / of the offset AND of its Jacobian travel together).  With a = act(z) the stored
            // value-row activation and t_c = act'(z) u_c the stored tangent rows:
            //   tangent rows:  u_bar_c = act'(z) t_bar_c
            //   value row   :  z_bar   = act'(z) a_bar + act''(z) sum_c t_bar_c u_c
            // and act'' u_c = 100 (1 - act') t_c for softplus(beta = 100), 0 for ReLU -- no division by act'.
            float d;
            if constexpr (ACT == SR_ACT_SOFTPLUS100) {
              if (use_stash) d = PF ? st[PF ? j : 0] : sg[e];   // the four rows of a point read the value row's stash entry
              else d = 1.0f - fast_ex2(kk * __shfl_sync(0xffffffffu, as, r.lane & ~3));
            } else if constexpr (ACT == SR_ACT_RELU) {
              d = __shfl_sync(0xffffffffu, as, r.lane & ~3) > 0.f ? 1.f : 0.f;
            } else d = 1.f;
            float cross = 0.f;
            if constexpr (ACT == SR_ACT_SOFTPLUS100) {
              // sum of t_bar_c * t_c over the three tangent lanes of the point (butterfly inside the lane quad)
              float prod = r.is_val ? 0.f : val * (as * a.mul_inv_scale);
              prod += __shfl_xor_sync(0xffffffffu, prod, 1);
              prod += __shfl_xor_sync(0xffffffffu, prod, 2);
              cross = prod;
            }
            if (c0 + j < a.n) {
              float o4 = val * d;
              if constexpr (ACT == SR_ACT_SOFTPLUS100) { if (r.is_val) o4 += 100.0f * (1.0f - d) * cross; }
              val = r.row_ok ? o4 : 0.f;
            }
          }
          o[j] = val * a.scale;
        }
      }
    } else {
      if (wait_v) tmem_wait(v);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + j4);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = j4 * 4 + jj;
          const float acc = __uint_as_float(v[j]);
          float val;
          if constexpr (CH == 1) {
            val = act_val<ACT>(acc + bb[jj]);
          } else {
            float d = 1.f;
            val = act_fn<ACT>(acc + bb[jj], d);                          // meaningful on value rows
            const float dv = __shfl_sync(0xffffffffu, d, r.lane & ~3);  // act'(z) of the value row
            if (!r.is_val) { val = dv * acc; }
            v[j] = __float_as_uint(d);
          }
          o[j] = val * a.scale;
        }
      }
      if (a.dstash != nullptr && r.is_val && r.row_ok) {
        if constexpr (CH == 1) {   // rarely used output: act' recomputed from the activation values
          const float inv = 1.0f / a.scale;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(dact_from_val<ACT>(o[j] * inv));
        }
        float4* dd = reinterpret_cast<float4*>(a.dstash + (size_t)r.row * r.ds_ld + c0);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          dd[j4] = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                               __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
      }
    }
    if (a.out != nullptr && r.row_ok && c0 < a.out_col0 + a.out_n && c0 + 32 > a.out_col0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = c0 + j - a.out_col0;
        if (c >= 0 && c < a.out_n) a.out[(size_t)r.row * a.out_ld + c] = o[j];
      }
    }
  }
  if (a.A_next == nullptr) return;
  const int kcn = c0 >> 5;  // next layer's k chunk
  if (kcn >= a.KCn) return;
  if (!live || c0 + 32 > a.n) {  // zero padding / skip-connection columns (uniform branch)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = c0 + j;
      if (!live || c >= a.n) {
        float val = 0.f;
        if (a.skip_src != nullptr && c >= a.n && c < a.n + a.skip_n && r.row_ok)
          val = a.skip_src[(size_t)r.row * a.skip_ld + (c - a.n)] * a.scale;
        o[j] = val;
      }
    }
  }
  __nv_bfloat16* base = a.A_next + a_tile_off(r.mt, kcn, a.KCn, 0) + (size_t)(r.row_in_tile >> 3) * 64 +
                        (r.row_in_tile & 7) * 8;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 q1, q2, q3;
    split3x2(o[g * 8 + 0], o[g * 8 + 1], q1.x, q2.x, q3.x);
    split3x2(o[g * 8 + 2], o[g * 8 + 3], q1.y, q2.y, q3.y);
    split3x2(o[g * 8 + 4], o[g * 8 + 5], q1.z, q2.z, q3.z);
    split3x2(o[g * 8 + 6], o[g * 8 + 7], q1.w, q2.w, q3.w);
    __nv_bfloat16* dst = base + (size_t)g * (BM * 8);
    *reinterpret_cast<uint4*>(dst) = q1;
    *reinterpret_cast<uint4*>(dst + A_PLANE) = q2;
    if constexpr (kPlanes == 3) *reinterpret_cast<uint4*>(dst + 2 * A_PLANE) = q3;
  }
}

template <int ACT, int CH, bool MUL>
__global__ void __launch_bounds__(epi_threads(MUL), 1) tc_layer_kernel(const __grid_constant__ LayerArgs a) {
  constexpr int kEpiWarps = epi_warps(MUL), kPartCols = epi_part_cols(kEpiWarps), kChunks = epi_chunks(kEpiWarps);
  constexpr bool kEpiDoubleBuffer = kEpiWarps == 8;
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)STAGES * A_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * (A_STAGE_BYTES + W_STAGE_BYTES));
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { sr_mbar_init(&full[i], 1); sr_mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sr_mbar_init(&tfull[i], 1); sr_mbar_init(&tempty[i], kEpiWarps); }
    sr_fence_barrier_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sr_smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long Mrows = a.M;
  int MTe = a.MT;
  if (a.m_dev != nullptr) {
    const long long md = (long long)(*a.m_dev);
    Mrows = md < a.M ? md : a.M;
    MTe = (int)((Mrows + BM - 1) / BM);
  }
  const long long ntiles = (long long)MTe * a.NT;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long mt = t / a.NT;
        const int nt = (int)(t % a.NT);
        for (int kc = 0; kc < a.KC; ++kc) {
          sr_mbar_wait(&empty[slot], phase ^ 1u);
#ifdef SR_TC_DBG_NOLOAD   // tuning knock-out (tools/tc_diag.py): no operand traffic
          sr_mbar_arrive(&full[slot]);
#else
          sr_mbar_arrive_expect_tx(&full[slot], A_STAGE_BYTES + W_STAGE_BYTES);
          sr_bulk_g2s(sA + (size_t)slot * A_STAGE, a.A + a_tile_off(mt, kc, a.KC, 0), A_STAGE_BYTES, &full[slot]);
          sr_bulk_g2s(sW + (size_t)slot * W_STAGE, a.W + w_tile_off(nt, kc, a.KC, 0), W_STAGE_BYTES, &full[slot]);
#endif
          if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      // plane pairs, smallest contributions first
      constexpr int kTerms = kPlanes == 2 ? 3 : 6;
      const int pa[6] = {0, 1, 0, 2, 1, 0};   // 2 planes: (a0,w1) (a1,w0) (a0,w0)
      const int pw[6] = {1, 0, 0, 0, 1, 0};   // 3 planes: see below
      const int pa3[6] = {0, 2, 1, 0, 1, 0};
      const int pw3[6] = {2, 0, 1, 1, 0, 0};
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int nt = (int)(t % a.NT);
        // N of this tile's MMAs: the GEMM's remaining columns rounded up to the instruction granule
        int mma_n = a.n_gemm - nt * BN;
        mma_n = mma_n >= BN ? BN : ((mma_n + 15) & ~15);
        const uint32_t idesc = kIdescBase | ((uint32_t)(mma_n >> 3) << 17);
        sr_mbar_wait(&tempty[buf], bphase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * BN;
        uint32_t accumulate = 0;
        for (int kc = 0; kc < a.KC; ++kc) {
          sr_mbar_wait(&full[slot], phase);
          tc_fence_after();
          const uint32_t abase = sr_smem_u32(sA + (size_t)slot * A_STAGE);
          const uint32_t wbase = sr_smem_u32(sW + (size_t)slot * W_STAGE);
#ifdef SR_TC_DBG_NOMMA    // tuning knock-out: 1/16 of the MMAs
          if (kc == 0)
#endif
#pragma unroll
          for (int q = 0; q < kTerms; ++q) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              // K = 16 per MMA = two 8-wide core matrices: advance two LBO steps per j
              const int qa = kPlanes == 2 ? pa[q] : pa3[q], qw = kPlanes == 2 ? pw[q] : pw3[q];
              const uint64_t ad = make_desc(abase + qa * (A_PLANE * 2) + j * 2 * (BM * 16), BM * 16, 128);
              const uint64_t bd = make_desc(wbase + qw * (W_PLANE * 2) + j * 2 * (BN * 16), BN * 16, 128);
              mma_bf16(tmem_d, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
          }
          mma_commit(&empty[slot]);  // frees the smem stage when these MMAs have read it
          if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        }
        mma_commit(&tfull[buf]);     // accumulator complete
        if (++buf == 2) { buf = 0; bphase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which half of the 256 accumulator columns
    EpiRow r;
    r.row_in_tile = q * 32 + lane;
    r.lane = lane;
    r.is_val = (CH == 1) || ((lane & 3) == 0);
    r.ds_ld = (size_t)a.NT * BN;
    int buf = 0;
    uint32_t bphase = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      r.mt = t / a.NT;
      r.nt = (int)(t % a.NT);
      r.row = r.mt * BM + r.row_in_tile;
      r.row_ok = r.row < Mrows;
      const int c_base = r.nt * BN + half * kPartCols;
#ifdef SR_TC_DBG_NOEPI    // tuning knock-out: accumulators are drained without being read
      const int n_live = 0;
      if (t >= 0) r.row_ok = false;
#else
      const int n_live = (a.n_gemm - c_base + 31) >> 5;   // chunks of this warp's half that hold GEMM columns
#endif
      sr_mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BN + half * kPartCols;
      if constexpr (CH == 1 && !MUL && kEpiDoubleBuffer) {
        // two register buffers: the TMEM load of chunk i+1 is in flight while chunk i is processed
        uint32_t va[32], vb[32];
        if (n_live > 0) tmem_ld32_async(taddr0, va);
#pragma unroll
        for (int i = 0; i < kChunks; i += 2) {
          if (i < n_live) tmem_wait(va);
          if (i + 1 < n_live) tmem_ld32_async(taddr0 + (i + 1) * 32, vb);
          epi_chunk<ACT, CH, MUL>(a, r, va, half * kChunks + i, i < n_live);
          if (i + 1 < n_live) tmem_wait(vb);
          if (i + 2 < kChunks && i + 2 < n_live) tmem_ld32_async(taddr0 + (i + 2) * 32, va);
          epi_chunk<ACT, CH, MUL>(a, r, vb, half * kChunks + i + 1, i + 1 < n_live);
        }
      } else {
        for (int i = 0; i < kChunks; ++i) {
          uint32_t v[32];
          // the chunk's global operands are requested inside epi_chunk BEFORE it waits for the TMEM load
          if (i < n_live) tmem_ld32_async(taddr0 + i * 32, v);
          epi_chunk<ACT, CH, MUL, kEpiWarps == 8>(a, r, v, half * kChunks + i, i < n_live, i < n_live);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) sr_mbar_arrive(&tempty[buf]);
      if (++buf == 2) { buf = 0; bphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---- CTA-pair variant (cta_group::2) ----------------------------------------------------------
// Two CTAs of a cluster (the two SMs of a TPC) share one 256 x 256 output tile: CTA r owns rows
// [128 r, 128 r + 128) (its own A tile and TMEM accumulator) and stages only HALF of the weight tile
// (columns [128 r, +128)); the leader (rank 0) issues `tcgen05.mma.cta_group::2` with M = 256, which reads
// A from each CTA's own shared memory and the two B halves from both.  Per SM that is 4 + 4 KB of operand
// reads per MMA instead of 4 + 8 and 32 KB instead of 48 KB of TMA fill per stage -- the single-CTA
// kernel is bound by exactly that shared-memory traffic (profiles/r01b_summary.md).
//   barriers (each CTA's own shared memory unless noted):
//     full[s]   own TMA bytes of stage s landed                    (producer expect_tx, count 1)
//     pfull[s]  LEADER only: the peer's stage s landed             (remote arrive by the peer's relay thread)
//     empty[s]  stage s consumed: leader's tcgen05.commit multicast to both CTAs
//     tfull[b]  accumulator b complete: leader's commit multicast to both CTAs
//     tempty[b] LEADER only: accumulator b drained by the epilogue warps of BOTH CTAs (2 x kEpiWarps)
constexpr int P_STAGES = 6;
constexpr int PB_PLANE = 128 * BK;                        // half weight tile plane: 128 rows x 32 k
constexpr int PB_STAGE = kPlanes * PB_PLANE;              // 16 KB (2 planes)
constexpr uint32_t PB_STAGE_BYTES = PB_STAGE * 2;
constexpr size_t kSmemPair = (size_t)P_STAGES * (A_STAGE_BYTES + PB_STAGE_BYTES) + 512;
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(sr_smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait on a local barrier that remote CTAs (or the async proxy of the pair) arrive on.  Default (CTA-scope)
// semantics on purpose: nothing written through the generic proxy by the partner is read here (operands
// arrive by TMA, accumulators through TMEM + tcgen05 fences), and a cluster-scope acquire compiles to
// MEMBAR.ALL.GPU + CCTL.IVALL in every poll -- measured 0.43 ms instead of 0.33 ms per layer.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_C:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE_C;\n"
      "bra.uni WAIT_LOOP_C;\n"
      "WAIT_DONE_C:\n"
      "}\n" ::"r"(sr_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          sr_smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <int ACT, int CH, bool MUL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(epi_threads(MUL), 1)
    tc_layer_pair_kernel(const __grid_constant__ LayerArgs a) {
  constexpr int kEpiWarps = epi_warps(MUL), kPartCols = epi_part_cols(kEpiWarps), kChunks = epi_chunks(kEpiWarps);
  constexpr bool kEpiDoubleBuffer = kEpiWarps == 8;
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)P_STAGES * A_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P_STAGES * (A_STAGE_BYTES + PB_STAGE_BYTES));
  uint64_t* full = bars;                    // [P_STAGES]
  uint64_t* pfull = bars + P_STAGES;        // [P_STAGES]  (used in the leader)
  uint64_t* empty = bars + 2 * P_STAGES;    // [P_STAGES]
  uint64_t* tfull = bars + 3 * P_STAGES;    // [2]
  uint64_t* tempty = tfull + 2;             // [2]         (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) { sr_mbar_init(&full[i], 1); sr_mbar_init(&pfull[i], 1); sr_mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sr_mbar_init(&tfull[i], 1); sr_mbar_init(&tempty[i], 2 * kEpiWarps); }
    sr_fence_barrier_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns in both CTAs (same warp id in both, cta_group::2)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sr_smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barrier inits of both CTAs visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long Mrows = a.M;
  int MTe = a.MT;
  if (a.m_dev != nullptr) {
    const long long md = (long long)(*a.m_dev);
    Mrows = md < a.M ? md : a.M;
    MTe = (int)((Mrows + BM - 1) / BM);
  }
  const long long npt = (long long)((MTe + 1) / 2) * a.NT;     // pair tiles (256 rows x 256 columns)
  const long long pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (long long t = pair_id; t < npt; t += npairs) {
        const long long mt = (t / a.NT) * 2 + rank;
        const int nt = (int)(t % a.NT);
        const bool have_a = mt < MTe;       // odd tile count: the peer's rows of the last pair do not exist
        for (int kc = 0; kc < a.KC; ++kc) {
          mbar_wait_cluster(&empty[slot], phase ^ 1u);
#ifdef SR_TC_DBG_NOLOAD
          sr_mbar_arrive(&full[slot]);
#else
          sr_mbar_arrive_expect_tx(&full[slot], (have_a ? A_STAGE_BYTES : 0u) + PB_STAGE_BYTES);
          if (have_a)
            sr_bulk_g2s(sA + (size_t)slot * A_STAGE, a.A + a_tile_off(mt, kc, a.KC, 0), A_STAGE_BYTES, &full[slot]);
          // this CTA's half of the weight tile (columns [128 rank, +128)): one contiguous block of the pair layout
          sr_bulk_g2s(sW + (size_t)slot * PB_STAGE,
                      a.Wp + (((size_t)nt * a.KC + kc) * 2 + rank) * PB_STAGE, PB_STAGE_BYTES, &full[slot]);
#endif
          if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      if (!leader) {
        // -------------------------------------------------------------- peer: relay "my stage landed"
        for (long long t = pair_id; t < npt; t += npairs) {
          for (int kc = 0; kc < a.KC; ++kc) {
            sr_mbar_wait(&full[slot], phase);
            mbar_arrive_remote(&pfull[slot], 0);
            if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
          }
        }
      } else {
        // -------------------------------------------------------------- leader: MMA issuer for the pair
        int buf = 0;
        uint32_t bphase = 0;
        constexpr int kTerms = kPlanes == 2 ? 3 : 6;
        const int pa[6] = {0, 1, 0, 2, 1, 0};
        const int pw[6] = {1, 0, 0, 0, 1, 0};
        const int pa3[6] = {0, 2, 1, 0, 1, 0};
        const int pw3[6] = {2, 0, 1, 1, 0, 0};
        for (long long t = pair_id; t < npt; t += npairs) {
          const int nt = (int)(t % a.NT);
          (void)nt;
          const uint32_t idesc = kIdescPair | ((uint32_t)(BN >> 3) << 17);
          mbar_wait_cluster(&tempty[buf], bphase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)buf * BN;
          uint32_t accumulate = 0;
          for (int kc = 0; kc < a.KC; ++kc) {
            sr_mbar_wait(&full[slot], phase);
#ifndef SR_TC_DBG_NOPFULL   // tuning knock-out: do not wait for the peer's stage
            mbar_wait_cluster(&pfull[slot], phase);
#endif
            tc_fence_after();
            const uint32_t abase = sr_smem_u32(sA + (size_t)slot * A_STAGE);
            const uint32_t wbase = sr_smem_u32(sW + (size_t)slot * PB_STAGE);
#ifdef SR_TC_DBG_NOMMA
            if (kc == 0)
#endif
#pragma unroll
            for (int q = 0; q < kTerms; ++q) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                const int qa = kPlanes == 2 ? pa[q] : pa3[q], qw = kPlanes == 2 ? pw[q] : pw3[q];
                const uint64_t ad = make_desc(abase + qa * (A_PLANE * 2) + j * 2 * (BM * 16), BM * 16, 128);
                const uint64_t bd = make_desc(wbase + qw * (PB_PLANE * 2) + j * 2 * (128 * 16), 128 * 16, 128);
                mma_bf16_pair(tmem_d, ad, bd, idesc, accumulate);
                accumulate = 1;
              }
            }
            mma_commit_pair(&empty[slot]);
            if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
          }
          mma_commit_pair(&tfull[buf]);
          if (++buf == 2) { buf = 0; bphase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    EpiRow r;
    r.row_in_tile = q * 32 + lane;
    r.lane = lane;
    r.is_val = (CH == 1) || ((lane & 3) == 0);
    r.ds_ld = (size_t)a.NT * BN;
    int buf = 0;
    uint32_t bphase = 0;
    for (long long t = pair_id; t < npt; t += npairs) {
      r.mt = (t / a.NT) * 2 + rank;
      r.nt = (int)(t % a.NT);
      r.row = r.mt * BM + r.row_in_tile;
      r.row_ok = r.row < Mrows;
#ifdef SR_TC_DBG_NOEPI
      const bool have_rows = false;
#else
      const bool have_rows = r.mt < MTe;
#endif
      const int c_base = r.nt * BN + half * kPartCols;
      const int n_live = have_rows ? (a.n_gemm - c_base + 31) >> 5 : 0;
      mbar_wait_cluster(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BN + half * kPartCols;
      if (!have_rows) {
        // phantom row tile of an odd tile count: nothing to read or write
      } else if constexpr (CH == 1 && !MUL && kEpiDoubleBuffer) {
        uint32_t va[32], vb[32];
        if (n_live > 0) tmem_ld32_async(taddr0, va);
#pragma unroll
        for (int i = 0; i < kChunks; i += 2) {
          if (i < n_live) tmem_wait(va);
          if (i + 1 < n_live) tmem_ld32_async(taddr0 + (i + 1) * 32, vb);
          epi_chunk<ACT, CH, MUL>(a, r, va, half * kChunks + i, i < n_live);
          if (i + 1 < n_live) tmem_wait(vb);
          if (i + 2 < kChunks && i + 2 < n_live) tmem_ld32_async(taddr0 + (i + 2) * 32, va);
          epi_chunk<ACT, CH, MUL>(a, r, vb, half * kChunks + i + 1, i + 1 < n_live);
        }
      } else {
        for (int i = 0; i < kChunks; ++i) {
          uint32_t v[32];
          // the chunk's global operands are requested inside epi_chunk BEFORE it waits for the TMEM load
          if (i < n_live) tmem_ld32_async(taddr0 + i * 32, v);
          epi_chunk<ACT, CH, MUL, kEpiWarps == 8>(a, r, v, half * kChunks + i, i < n_live, i < n_live);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty[buf], 0);   // the leader's barrier counts both CTAs' warps
      if (++buf == 2) { buf = 0; bphase ^= 1u; }
    }
  }
  tc_fence_before();
  cluster_sync_all();      // no CTA leaves (or frees TMEM) while its partner may still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- whole-sweep kernel: all layers of a forward or reverse sweep in ONE launch ---------------------------
// A CTA pair keeps its row tiles through every layer of the sweep: pair p owns the row-pair tiles
// {p, p + npairs, ...} (256 rows each) in EVERY step.  The activations a step's epilogue writes for rows R are
// exactly what the same CTA loads as the next step's A operand for rows R, so the only dependency between steps
// is inside a CTA: the producer waits until the eight epilogue warps have finished the previous step on that row
// tile (per-warp progress counters in shared memory, release / acquire, with a generic->async proxy fence on both
// sides because the tiles are written with st.global and read back by the bulk-copy engine).  No grid or cluster
// barrier between layers, one TMEM allocation, one barrier set-up and one launch ramp per SWEEP instead of per
// layer; the smem ring and the two TMEM accumulators simply keep rolling across the layer boundary, so the first
// MMAs of step l+1 overlap the last epilogue of step l whenever a pair owns more than one row tile.
// Narrow steps (N < 256: the SDF's last layer, the input gradient of the reverse sweep) run as one N = 256 tile on
// zero-padded weight rows.  Buffers that hold tiles of different widths must be distinct: tile (mt, kc) sits at
// (mt * KC + kc), so two layouts of one buffer alias ACROSS row tiles (the host wrapper checks this).
constexpr int kMaxSteps = 12;
static_assert(kPlanes == 2, "the whole-sweep kernel issues the 3-term (2-plane) product");
struct StepArgs {
  LayerArgs la;
  int act;       // SR_ACT_* of a forward step / of the PREVIOUS layer for a reverse (mul) step
  int mul;       // 1: reverse step (acc * act'(z_prev))
};
struct SweepArgs {
  StepArgs step[kMaxSteps];
  int L;
  int dbg;          // tuning knock-outs (sr_tc_debug_sweep_flags): 1 = no proxy fence in the epilogue (WRONG results)
  int last_plain;   // the last step is a plain linear step (no activation, no act' multiply): the network's output layer
                    // in a forward sweep, the input gradient in a reverse sweep
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_release_smem(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(sr_smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_smem(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sr_smem_u32(p)) : "memory");
  return v;
}

// one accumulator (this warp's 32 lanes x its 128-column half) -> epilogue
template <int ACT, int CH, bool MUL, int EW, bool DB = true>
__device__ __forceinline__ void epi_item(const LayerArgs& a, const EpiRow& r, uint32_t taddr0, int n_live, int half) {
  constexpr int kChunks = epi_chunks(EW);
  if constexpr (CH == 1 && !MUL && DB && EW == 8) {
    uint32_t va[32], vb[32];
    if (n_live > 0) tmem_ld32_async(taddr0, va);
#pragma unroll
    for (int i = 0; i < kChunks; i += 2) {
      if (i < n_live) tmem_wait(va);
      if (i + 1 < n_live) tmem_ld32_async(taddr0 + (i + 1) * 32, vb);
      epi_chunk<ACT, CH, MUL>(a, r, va, half * kChunks + i, i < n_live);
      if (i + 1 < n_live) tmem_wait(vb);
      if (i + 2 < kChunks && i + 2 < n_live) tmem_ld32_async(taddr0 + (i + 2) * 32, va);
      epi_chunk<ACT, CH, MUL>(a, r, vb, half * kChunks + i + 1, i + 1 < n_live);
    }
  } else {
    for (int i = 0; i < kChunks; ++i) {
      uint32_t v[32];
      if (i < n_live) tmem_ld32_async(taddr0 + i * 32, v);
      epi_chunk<ACT, CH, MUL, EW == 8>(a, r, v, half * kChunks + i, i < n_live, i < n_live);
    }
  }
}

// <ACT, MUL>: activation / mode of every step but (optionally) the last -- the two epilogues a sweep needs
template <int ACT, int CH, bool MUL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(epi_threads(MUL), 1)
    tc_sweep_pair_kernel(const __grid_constant__ SweepArgs sw) {
  constexpr int kEpiWarps = epi_warps(MUL), kPartCols = epi_part_cols(kEpiWarps);
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)P_STAGES * A_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P_STAGES * (A_STAGE_BYTES + PB_STAGE_BYTES));
  uint64_t* full = bars;                    // [P_STAGES]
  uint64_t* pfull = bars + P_STAGES;        // [P_STAGES]  (used in the leader)
  uint64_t* empty = bars + 2 * P_STAGES;    // [P_STAGES]
  uint64_t* tfull = bars + 3 * P_STAGES;    // [2]
  uint64_t* tempty = tfull + 2;             // [2]         (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* done = tmem_slot + 2;           // [kEpiWarps] items finished by each epilogue warp of THIS CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) { sr_mbar_init(&full[i], 1); sr_mbar_init(&pfull[i], 1); sr_mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sr_mbar_init(&tfull[i], 1); sr_mbar_init(&tempty[i], 2 * kEpiWarps); }
    for (int i = 0; i < kEpiWarps; ++i) done[i] = 0u;
    sr_fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sr_smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();         // `done` zeroed before any role reads it
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const LayerArgs& a0 = sw.step[0].la;
  long long Mrows = a0.M;
  int MTe = a0.MT;
  if (a0.m_dev != nullptr) {
    const long long md = (long long)(*a0.m_dev);
    Mrows = md < a0.M ? md : a0.M;
    MTe = (int)((Mrows + BM - 1) / BM);
  }
  const int nrp = (MTe + 1) / 2;                                   // row-pair tiles (256 rows)
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int J = pair_id < nrp ? (nrp - pair_id + npairs - 1) / npairs : 0;   // row-pair tiles of this pair
  const int L = sw.L;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      uint32_t items_before = 0;            // items of the steps before the previous one
      for (int l = 0; l < L; ++l) {
        const LayerArgs& a = sw.step[l].la;
        const int NTp = l > 0 ? sw.step[l - 1].la.NT : 0;
        for (int j = 0; j < J; ++j) {
          const long long mt = (long long)(pair_id + j * npairs) * 2 + rank;
          const bool have_a = mt < MTe;
          for (int nt = 0; nt < a.NT; ++nt) {
            for (int kc = 0; kc < a.KC; ++kc) {
              mbar_wait_cluster(&empty[slot], phase ^ 1u);
              sr_mbar_arrive_expect_tx(&full[slot], (have_a ? A_STAGE_BYTES : 0u) + PB_STAGE_BYTES);
              sr_bulk_g2s(sW + (size_t)slot * PB_STAGE,
                          a.Wp + (((size_t)nt * a.KC + kc) * 2 + rank) * PB_STAGE, PB_STAGE_BYTES, &full[slot]);
              if (l > 0 && nt == 0 && kc == 0) {
                // rows of tile j: every n-tile of the previous step must have left the epilogue (this CTA's warps)
                const uint32_t need = items_before + (uint32_t)(j + 1) * (uint32_t)NTp;
                for (int w = 0; w < kEpiWarps; ++w)
                  while (ld_acquire_smem(&done[w]) < need) {}
                fence_proxy_async();
              }
              if (have_a)
                sr_bulk_g2s(sA + (size_t)slot * A_STAGE, a.A + a_tile_off(mt, kc, a.KC, 0), A_STAGE_BYTES, &full[slot]);
              if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
            }
          }
        }
        if (l > 0) items_before += (uint32_t)J * (uint32_t)NTp;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      if (!leader) {
        for (int l = 0; l < L; ++l) {
          const LayerArgs& a = sw.step[l].la;
          const int n_it = J * a.NT * a.KC;
          for (int it = 0; it < n_it; ++it) {
            sr_mbar_wait(&full[slot], phase);
            mbar_arrive_remote(&pfull[slot], 0);
            if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
          }
        }
      } else {
        int buf = 0;
        uint32_t bphase = 0;
        const int pa[3] = {0, 1, 0};
        const int pw[3] = {1, 0, 0};
        const uint32_t idesc = kIdescPair | ((uint32_t)(BN >> 3) << 17);
        for (int l = 0; l < L; ++l) {
          const LayerArgs& a = sw.step[l].la;
          const int n_items = J * a.NT;
          for (int item = 0; item < n_items; ++item) {
            mbar_wait_cluster(&tempty[buf], bphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * BN;
            uint32_t accumulate = 0;
            for (int kc = 0; kc < a.KC; ++kc) {
              sr_mbar_wait(&full[slot], phase);
              mbar_wait_cluster(&pfull[slot], phase);
              tc_fence_after();
              const uint32_t abase = sr_smem_u32(sA + (size_t)slot * A_STAGE);
              const uint32_t wbase = sr_smem_u32(sW + (size_t)slot * PB_STAGE);
#pragma unroll
              for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int jj = 0; jj < BK / 16; ++jj) {
                  const uint64_t ad = make_desc(abase + pa[q] * (A_PLANE * 2) + jj * 2 * (BM * 16), BM * 16, 128);
                  const uint64_t bd = make_desc(wbase + pw[q] * (PB_PLANE * 2) + jj * 2 * (128 * 16), 128 * 16, 128);
                  mma_bf16_pair(tmem_d, ad, bd, idesc, accumulate);
                  accumulate = 1;
                }
              }
              mma_commit_pair(&empty[slot]);
              if (++slot == P_STAGES) { slot = 0; phase ^= 1u; }
            }
            mma_commit_pair(&tfull[buf]);
            if (++buf == 2) { buf = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    EpiRow r;
    r.row_in_tile = q * 32 + lane;
    r.lane = lane;
    r.is_val = (CH == 1) || ((lane & 3) == 0);
    int buf = 0;
    uint32_t bphase = 0;
    uint32_t items = 0;
    for (int l = 0; l < L; ++l) {
      const LayerArgs& a = sw.step[l].la;
      const bool plain = sw.last_plain && l == L - 1;
      r.ds_ld = (size_t)a.NT * BN;
      for (int j = 0; j < J; ++j) {
        r.mt = (long long)(pair_id + j * npairs) * 2 + rank;
        r.row = r.mt * BM + r.row_in_tile;
        r.row_ok = r.row < Mrows;
        const bool have_rows = r.mt < MTe;
        for (int nt = 0; nt < a.NT; ++nt) {
          r.nt = nt;
          const int c_base = nt * BN + half * kPartCols;
          const int n_live = have_rows ? (a.n_gemm - c_base + 31) >> 5 : 0;
          mbar_wait_cluster(&tfull[buf], bphase);
          tc_fence_after();
          const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BN + half * kPartCols;
          if (have_rows) {
            if (plain) epi_item<SR_ACT_NONE, CH, false, kEpiWarps, false>(a, r, taddr0, n_live, half);
            else epi_item<ACT, CH, MUL, kEpiWarps>(a, r, taddr0, n_live, half);
          }
          tc_fence_before();
          // this lane's tile stores -> visible to the bulk copies of the next step (nothing reads the last step's)
          // (one fence + one progress update per ROW TILE: the next step needs all n-tiles of the row tile anyway;
          //  fence.proxy.async is MEMBAR.GPU + FENCE.VIEW.ASYNC, i.e. it waits for this lane's stores)
          ++items;
          const bool publish = nt == a.NT - 1 && l + 1 < L;
          if (publish && !(sw.dbg & 1)) fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_remote(&tempty[buf], 0);
            if (publish) st_release_smem(&done[warp - 2], items);
          }
          if (++buf == 2) { buf = 0; bphase ^= 1u; }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- packing kernels -------------------------------------------------------------------------
// fp32 row-major [M][K] (ld) -> tiled split-bf16 activations with KC = ceil(Kpad/32) chunks
__global__ void pack_rows_kernel(const float* __restrict__ src, long long M, int K, int ld,
                                 __nv_bfloat16* __restrict__ dst, int KC, long long MT,
                                 const int* __restrict__ m_dev) {
  if (m_dev != nullptr) {
    const long long md = *m_dev;
    M = md < M ? md : M;
    MT = (M + BM - 1) / BM;
  }
  const long long total = MT * BM * (long long)KC * 4;  // one thread per (row, k8 group)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % BM);
    const long long rest = idx / BM;
    const int g = (int)(rest % (KC * 4));
    const long long mt = rest / (KC * 4);
    const long long row = mt * BM + r;
    const int kc = g >> 2, k8 = g & 3;
    __align__(16) __nv_bfloat16 p1[8], p2[8], p3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 32 + k8 * 8 + e;
      const float x = (row < M && k < K) ? src[(size_t)row * ld + k] : 0.f;
      split3(x, p1[e], p2[e], p3[e]);
    }
    const size_t off = (size_t)k8 * (BM * 8) + (size_t)(r >> 3) * 64 + (r & 7) * 8;
    *reinterpret_cast<uint4*>(dst + a_tile_off(mt, kc, KC, 0) + off) = *reinterpret_cast<uint4*>(p1);
    *reinterpret_cast<uint4*>(dst + a_tile_off(mt, kc, KC, 1) + off) = *reinterpret_cast<uint4*>(p2);
    if constexpr (kPlanes == 3)
      *reinterpret_cast<uint4*>(dst + a_tile_off(mt, kc, KC, kPlanes - 1) + off) = *reinterpret_cast<uint4*>(p3);
  }
}

// effective weights, fp32 row-major [N][K] (ld) -> tiled split-bf16 [NT][KC][planes][256x32] followed by the CTA-pair layout
__global__ void pack_weights_kernel(const float* __restrict__ w, int N, int K, int ld,
                                    __nv_bfloat16* __restrict__ dst, int NT, int KC) {
  const long long total = (long long)NT * BN * KC * 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % BN);
    const long long rest = idx / BN;
    const int g = (int)(rest % (KC * 4));
    const int nt = (int)(rest / (KC * 4));
    const int n = nt * BN + r;
    const int kc = g >> 2, k8 = g & 3;
    __align__(16) __nv_bfloat16 p1[8], p2[8], p3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 32 + k8 * 8 + e;
      const float x = (n < N && k < K) ? w[(size_t)n * ld + k] : 0.f;
      split3(x, p1[e], p2[e], p3[e]);
    }
    const size_t off = (size_t)k8 * (BN * 8) + (size_t)(r >> 3) * 64 + (r & 7) * 8;
    *reinterpret_cast<uint4*>(dst + w_tile_off(nt, kc, KC, 0) + off) = *reinterpret_cast<uint4*>(p1);
    *reinterpret_cast<uint4*>(dst + w_tile_off(nt, kc, KC, 1) + off) = *reinterpret_cast<uint4*>(p2);
    // second copy in the CTA-pair layout: [nt][kc][half][plane][k8][128 rows][8]
    __nv_bfloat16* dp = dst + (size_t)NT * KC * W_STAGE + (((size_t)nt * KC + kc) * 2 + (r >> 7)) * PB_STAGE +
                        (size_t)k8 * (128 * 8) + (size_t)((r & 127) >> 3) * 64 + (r & 7) * 8;
    *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<uint4*>(p1);
    *reinterpret_cast<uint4*>(dp + PB_PLANE) = *reinterpret_cast<uint4*>(p2);
    if constexpr (kPlanes == 3) *reinterpret_cast<uint4*>(dp + 2 * PB_PLANE) = *reinterpret_cast<uint4*>(p3);
    if constexpr (kPlanes == 3)
      *reinterpret_cast<uint4*>(dst + w_tile_off(nt, kc, KC, kPlanes - 1) + off) = *reinterpret_cast<uint4*>(p3);
  }
}

// Embedded network input, fp32 row-major [P*ch][ld]: PE(p) (+ cond[b]) for value rows and
// d/dp_t of it for tangent rows (model/Embedder.py:11-32).  One thread per element (coalesced).
struct EmbedArgs {
  const float* pts;
  long long P;
  int multires;
  float pe_w[16];
  int ch;
  const float* conds;
  const long long* batch_inds;
  long long pts_per_frame;
  int condlen;
  float* out;
  int ld;
  const int* index;   // optional active list: row i embeds point index[i]
  const int* m_dev;   // optional device-side count of active points
};
__global__ void embed_kernel(const __grid_constant__ EmbedArgs a) {
  long long np = a.P;
  if (a.m_dev != nullptr) { const long long md = *a.m_dev; np = md < np ? md : np; }
  const long long total = np * a.ch * (long long)a.ld;
  const int pe_dim = 3 + 6 * a.multires;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % a.ld);
    const long long row = idx / a.ld;
    const long long pi = row / a.ch;
    const long long p = a.index ? (long long)a.index[pi] : pi;
    const int t = (int)(row % a.ch);  // 0 = value, 1..3 = d/dp_{t-1}
    float v = 0.f;
    if (k < 3) {
      v = t == 0 ? a.pts[p * 3 + k] : (t - 1 == k ? 1.f : 0.f);
    } else if (k < pe_dim) {
      const int b = (k - 3) / 6, w6 = (k - 3) % 6, j = w6 % 3;
      const bool is_cos = w6 >= 3;
      if (t == 0 || t - 1 == j) {
        const float freq = (float)(1 << b);
        float sn, cs;
        sincosf(a.pts[p * 3 + j] * freq, &sn, &cs);
        const float w = a.pe_w[b];
        if (t == 0) v = w * (is_cos ? cs : sn);
        else v = is_cos ? -(w * freq) * sn : (w * freq) * cs;
      }
    } else if (k < pe_dim + a.condlen) {
      if (t == 0) {
        const long long bi = a.batch_inds ? a.batch_inds[p] : (a.pts_per_frame > 0 ? p / a.pts_per_frame : 0);
        v = a.conds[bi * a.condlen + (k - pe_dim)];
      }
    }
    a.out[idx] = v;
  }
}

// Backward of embed_kernel w.r.t. the points: gx [P*ch][ld] cotangent rows -> gp [P][3].  Value row: d PE / d p; tangent
// rows (ch = 4): the tangent entries themselves depend on p (second derivative of the encoding).  One thread per point.
__global__ void embed_bwd_kernel(const float* __restrict__ pts, long long P, int multires, const float* __restrict__ gx,
                                 int ld, int ch, float* __restrict__ gp, float pw0, float pw1, float pw2, float pw3,
                                 float pw4, float pw5, float pw6, float pw7) {
  const float pw[8] = {pw0, pw1, pw2, pw3, pw4, pw5, pw6, pw7};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const float* gv = gx + (size_t)p * ch * ld;
    float out[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float x = pts[p * 3 + j];
      float acc = gv[j];
      const float* gt = ch == 4 ? gv + (size_t)(1 + j) * ld : nullptr;   // tangent row d/dp_j: only column j is non-zero
      float freq = 1.0f;
      for (int b = 0; b < multires; ++b, freq *= 2.0f) {
        float sn, cs;
        sincosf(x * freq, &sn, &cs);
        const float w = pw[b] * freq;
        const int ks = 3 + 6 * b + j, kc = ks + 3;
        acc += w * (cs * gv[ks] - sn * gv[kc]);
        if (gt) acc -= w * freq * (sn * gt[ks] + cs * gt[kc]);
      }
      out[j] = acc;
    }
    gp[p * 3] = out[0]; gp[p * 3 + 1] = out[1]; gp[p * 3 + 2] = out[2];
  }
}

}  // namespace sr_tc

extern "C" {

int sr_tc_embed_backward(const float* pts, int64_t P, int multires, const float* pe_w, int ch, const float* gx, int ld,
                         float* gp, cudaStream_t s) {
  if (!pts || !gx || !gp || !pe_w || P <= 0 || (ch != 1 && ch != 4) || multires < 0 || multires > 8 ||
      ld < 3 + 6 * multires)
    return SR_EINVAL;
  float w[8];
  for (int i = 0; i < 8; ++i) w[i] = i < multires ? pe_w[i] : 0.f;
  sr_tc::embed_bwd_kernel<<<sr_grid_for(P, 256, 8), 256, 0, s>>>(pts, P, multires, gx, ld, ch, gp, w[0], w[1], w[2], w[3],
                                                                  w[4], w[5], w[6], w[7]);
  return sr_launch_status();
}

int sr_tc_embed(const float* pts, int64_t P, int multires, const float* pe_w, int ch,
                const float* conds, const int64_t* batch_inds, int64_t pts_per_frame, int condlen,
                float* out, int ld, const int32_t* index, const int32_t* m_dev, cudaStream_t s) {
  if (!pts || !out || !pe_w || P <= 0 || (ch != 1 && ch != 4) || multires < 0 || multires > 16) return SR_EINVAL;
  if (ld < 3 + 6 * multires + condlen || (condlen > 0 && !conds)) return SR_EINVAL;
  sr_tc::EmbedArgs a;
  a.pts = pts; a.P = P; a.multires = multires; a.ch = ch; a.conds = conds;
  a.batch_inds = (const long long*)batch_inds; a.pts_per_frame = pts_per_frame; a.condlen = condlen;
  a.out = out; a.ld = ld; a.index = index; a.m_dev = m_dev;
  for (int i = 0; i < 16; ++i) a.pe_w[i] = i < multires ? pe_w[i] : 0.f;
  const long long total = P * ch * (long long)ld;
  sr_tc::embed_kernel<<<sr_grid_for(total, 256, 8), 256, 0, s>>>(a);
  return sr_launch_status();
}

int64_t sr_tc_act_bytes(int64_t M, int K) {
  const int64_t MT = (M + sr_tc::BM - 1) / sr_tc::BM, KC = (K + 31) / 32;
  return MT * KC * sr_tc::kPlanes * sr_tc::A_PLANE * 2;
}
int64_t sr_tc_weight_bytes(int N, int K) {
  const int64_t NT = (N + sr_tc::BN - 1) / sr_tc::BN, KC = (K + 31) / 32;
  return 2 * NT * KC * sr_tc::kPlanes * sr_tc::W_PLANE * 2;   // single-CTA layout + CTA-pair layout
}

int sr_tc_pack_rows(const float* src, int64_t M, int K, int ld, void* dst, const int32_t* m_dev,
                    cudaStream_t s) {
  if (!src || !dst || M <= 0 || K <= 0 || ld < K) return SR_EINVAL;
  const long long MT = (M + sr_tc::BM - 1) / sr_tc::BM;
  const int KC = (K + 31) / 32;
  const long long total = MT * sr_tc::BM * KC * 4;
  sr_tc::pack_rows_kernel<<<sr_grid_for(total, 256, 8), 256, 0, s>>>(src, M, K, ld, (__nv_bfloat16*)dst, KC, MT, m_dev);
  return sr_launch_status();
}

int sr_tc_pack_weights(const float* w, int N, int K, int ld, void* dst, cudaStream_t s) {
  if (!w || !dst || N <= 0 || K <= 0 || ld < K) return SR_EINVAL;
  const int NT = (N + sr_tc::BN - 1) / sr_tc::BN, KC = (K + 31) / 32;
  const long long total = (long long)NT * sr_tc::BN * KC * 4;
  sr_tc::pack_weights_kernel<<<sr_grid_for(total, 256, 8), 256, 0, s>>>(w, N, K, ld, (__nv_bfloat16*)dst, NT, KC);
  return sr_launch_status();
}

int sr_tc_linear(const void* A, const void* W, const float* bias, int64_t M, int N, int K, int n_valid,
                 int act, int ch, void* A_next, int K_next, float scale, const float* skip_src,
                 int skip_n, int skip_ld, float* out, int out_ld, int out_col0, int out_n,
                 float* dstash, const void* mul_tiles, int mul_K, int mul_act, float mul_scale,
                 const int32_t* m_dev, cudaStream_t s) {
  using namespace sr_tc;
  if (!A || !W || !bias || M <= 0 || N <= 0 || K <= 0 || (ch != 1 && ch != 4)) return SR_EINVAL;
  if (!A_next && !out) return SR_EINVAL;
  LayerArgs a;
  a.A = (const __nv_bfloat16*)A; a.W = (const __nv_bfloat16*)W; a.bias = bias; a.M = M;
  a.MT = (int)((M + BM - 1) / BM); a.NT = (N + BN - 1) / BN; a.KC = (K + 31) / 32;
  a.Wp = a.W + (size_t)a.NT * a.KC * W_STAGE;
  a.n_gemm = N; a.n = n_valid; a.ch = ch;
  a.A_next = (__nv_bfloat16*)A_next; a.KCn = A_next ? (K_next + 31) / 32 : 0;
  a.scale = scale; a.skip_src = skip_src; a.skip_n = skip_n; a.skip_ld = skip_ld;
  a.out = out; a.out_ld = out_ld; a.dstash = dstash; a.out_col0 = out_col0; a.out_n = out_n;
  a.mul_tiles = (const __nv_bfloat16*)mul_tiles; a.mul_KC = (mul_K + 31) / 32;
  a.mul_inv_scale = mul_scale != 0.f ? 1.0f / mul_scale : 1.0f; a.m_dev = m_dev;
  if (mul_tiles && mul_K < n_valid) return SR_EINVAL;
  using Kern = void (*)(const LayerArgs);
  static const int use_pair = [] { const char* e = getenv("SELFRECON_B200_TC_PAIR"); return e ? atoi(e) : 1; }();
  // the pair kernel always issues N = 256 MMAs (a narrower N would take columns from BOTH halves)
  const bool pair = use_pair && a.MT >= 2 && (N % BN == 0);
#define SR_TC_PICK(ACT_, CH_, MUL_) \
  (pair ? (Kern)tc_layer_pair_kernel<ACT_, CH_, MUL_> : (Kern)tc_layer_kernel<ACT_, CH_, MUL_>)
  Kern kern = nullptr;
  if (mul_tiles && ch == 4) {
    switch (mul_act) {
      case SR_ACT_NONE: kern = SR_TC_PICK(SR_ACT_NONE, 4, true); break;
      case SR_ACT_SOFTPLUS100: kern = SR_TC_PICK(SR_ACT_SOFTPLUS100, 4, true); break;
      case SR_ACT_RELU: kern = SR_TC_PICK(SR_ACT_RELU, 4, true); break;
    }
  } else if (mul_tiles) {
    switch (mul_act) {
      case SR_ACT_NONE: kern = SR_TC_PICK(SR_ACT_NONE, 1, true); break;
      case SR_ACT_SOFTPLUS100: kern = SR_TC_PICK(SR_ACT_SOFTPLUS100, 1, true); break;
      case SR_ACT_RELU: kern = SR_TC_PICK(SR_ACT_RELU, 1, true); break;
    }
  } else if (ch == 1) {
    switch (act) {
      case SR_ACT_NONE: kern = SR_TC_PICK(SR_ACT_NONE, 1, false); break;
      case SR_ACT_SOFTPLUS100: kern = SR_TC_PICK(SR_ACT_SOFTPLUS100, 1, false); break;
      case SR_ACT_RELU: kern = SR_TC_PICK(SR_ACT_RELU, 1, false); break;
      case SR_ACT_TANH: kern = SR_TC_PICK(SR_ACT_TANH, 1, false); break;
    }
  } else {
    switch (act) {
      case SR_ACT_NONE: kern = SR_TC_PICK(SR_ACT_NONE, 4, false); break;
      case SR_ACT_SOFTPLUS100: kern = SR_TC_PICK(SR_ACT_SOFTPLUS100, 4, false); break;
      case SR_ACT_RELU: kern = SR_TC_PICK(SR_ACT_RELU, 4, false); break;
      case SR_ACT_TANH: kern = SR_TC_PICK(SR_ACT_TANH, 4, false); break;
    }
  }
#undef SR_TC_PICK
  if (!kern) return SR_EINVAL;
  // cudaFuncSetAttribute is per device: one flag per device ordinal (a process may drive several GPUs)
  static bool attr_set_dev[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
#define SR_TC_BOTH(ACT_, CH_, MUL_)                                                                        \
  {                                                                                                        \
    cudaError_t e1 = cudaFuncSetAttribute(tc_layer_kernel<ACT_, CH_, MUL_>,                                \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);        \
    cudaError_t e2 = cudaFuncSetAttribute(tc_layer_pair_kernel<ACT_, CH_, MUL_>,                           \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPair);    \
    if (e1 != cudaSuccess) return (int)e1;                                                                 \
    if (e2 != cudaSuccess) return (int)e2;                                                                 \
  }
    SR_TC_BOTH(SR_ACT_NONE, 1, true) SR_TC_BOTH(SR_ACT_SOFTPLUS100, 1, true) SR_TC_BOTH(SR_ACT_RELU, 1, true)
    SR_TC_BOTH(SR_ACT_NONE, 4, true) SR_TC_BOTH(SR_ACT_SOFTPLUS100, 4, true) SR_TC_BOTH(SR_ACT_RELU, 4, true)
    SR_TC_BOTH(SR_ACT_NONE, 1, false) SR_TC_BOTH(SR_ACT_SOFTPLUS100, 1, false) SR_TC_BOTH(SR_ACT_RELU, 1, false)
    SR_TC_BOTH(SR_ACT_TANH, 1, false) SR_TC_BOTH(SR_ACT_NONE, 4, false) SR_TC_BOTH(SR_ACT_SOFTPLUS100, 4, false)
    SR_TC_BOTH(SR_ACT_RELU, 4, false) SR_TC_BOTH(SR_ACT_TANH, 4, false)
#undef SR_TC_BOTH
    attr_set = true;
  }
  if (pair) {
    const long long npt = (long long)((a.MT + 1) / 2) * a.NT;
    const long long npairs = npt < SR_NUM_SMS_B200 / 2 ? npt : SR_NUM_SMS_B200 / 2;
    kern<<<(unsigned)(2 * npairs), epi_threads(mul_tiles != nullptr), kSmemPair, s>>>(a);
  } else {
    const long long ntiles = (long long)a.MT * a.NT;
    const int grid = (int)(ntiles < SR_NUM_SMS_B200 ? ntiles : SR_NUM_SMS_B200);
    kern<<<grid, epi_threads(mul_tiles != nullptr), kSmem, s>>>(a);
  }
  return sr_launch_status();
}
static int g_sweep_dbg = 0;
int sr_tc_debug_sweep_flags(int flags) {
  const int old = g_sweep_dbg;
  g_sweep_dbg = flags;
  return old;
}

int sr_tc_sweep(const sr_tc_step* steps, int L, int64_t M, int ch, const int32_t* m_dev, cudaStream_t s) {
  using namespace sr_tc;
  if (!steps || L <= 0 || L > kMaxSteps || M <= 0 || (ch != 1 && ch != 4)) return SR_EINVAL;
  SweepArgs sw;
  sw.L = L;
  sw.dbg = g_sweep_dbg;
  const int MT = (int)((M + BM - 1) / BM);
  for (int l = 0; l < L; ++l) {
    const sr_tc_step& t = steps[l];
    if (!t.A || !t.W || !t.bias || t.N <= 0 || t.K <= 0 || (!t.A_next && !t.out)) return SR_EINVAL;
    if (t.act != SR_ACT_NONE && t.act != SR_ACT_SOFTPLUS100 && t.act != SR_ACT_RELU) return SR_EINVAL;
    if (t.mul_tiles && (t.mul_act != SR_ACT_NONE && t.mul_act != SR_ACT_SOFTPLUS100 && t.mul_act != SR_ACT_RELU))
      return SR_EINVAL;
    if (t.mul_tiles && t.mul_K < t.n_valid) return SR_EINVAL;
    LayerArgs& a = sw.step[l].la;
    a.A = (const __nv_bfloat16*)t.A; a.W = (const __nv_bfloat16*)t.W; a.bias = t.bias; a.M = M;
    a.MT = MT; a.NT = (t.N + BN - 1) / BN; a.KC = (t.K + 31) / 32;
    a.Wp = a.W + (size_t)a.NT * a.KC * W_STAGE;
    a.n_gemm = t.N; a.n = t.n_valid; a.ch = ch;
    a.A_next = (__nv_bfloat16*)t.A_next; a.KCn = t.A_next ? (t.K_next + 31) / 32 : 0;
    a.scale = t.scale; a.skip_src = t.skip_src; a.skip_n = t.skip_n; a.skip_ld = t.skip_ld;
    a.out = t.out; a.out_ld = t.out_ld; a.dstash = t.dstash; a.out_col0 = t.out_col0; a.out_n = t.out_n;
    a.mul_tiles = (const __nv_bfloat16*)t.mul_tiles; a.mul_KC = (t.mul_K + 31) / 32;
    a.mul_inv_scale = t.mul_scale != 0.f ? 1.0f / t.mul_scale : 1.0f; a.m_dev = m_dev;
    sw.step[l].mul = t.mul_tiles ? 1 : 0;
    sw.step[l].act = t.mul_tiles ? t.mul_act : t.act;
  }
  // one buffer, two tile widths: the layouts alias across row tiles, and pairs run through the steps unsynchronised
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const LayerArgs &x = sw.step[i].la, &y = sw.step[j].la;
      if (x.A_next && x.A_next == y.A_next && x.KCn != y.KCn) return SR_EINVAL;
      if (x.A_next && (const void*)x.A_next == (const void*)y.A && x.KCn != y.KC) return SR_EINVAL;
      if (x.A == y.A && x.KC != y.KC) return SR_EINVAL;
    }
  // every step but an optional plain last one shares (activation, mode)
  const int body_act = sw.step[0].act, body_mul = sw.step[0].mul;
  const StepArgs& lastp = sw.step[L - 1];
  sw.last_plain = (L > 1 || true) && !lastp.mul && lastp.act == SR_ACT_NONE && (body_mul || body_act != SR_ACT_NONE) ? 1 : 0;
  for (int l = 0; l < L - (sw.last_plain ? 1 : 0); ++l)
    if (sw.step[l].act != body_act || sw.step[l].mul != body_mul) return SR_EINVAL;
  using Kern = void (*)(const SweepArgs);
  Kern kern = nullptr;
#define SR_SW(ACT_, CH_, MUL_) kern = (Kern)tc_sweep_pair_kernel<ACT_, CH_, MUL_>
  const int key = (ch == 4 ? 8 : 0) + (body_mul ? 4 : 0) + body_act;
  switch (key) {
    case 0: SR_SW(SR_ACT_NONE, 1, false); break;
    case 1: SR_SW(SR_ACT_SOFTPLUS100, 1, false); break;
    case 2: SR_SW(SR_ACT_RELU, 1, false); break;
    case 4: SR_SW(SR_ACT_NONE, 1, true); break;
    case 5: SR_SW(SR_ACT_SOFTPLUS100, 1, true); break;
    case 6: SR_SW(SR_ACT_RELU, 1, true); break;
    case 8: SR_SW(SR_ACT_NONE, 4, false); break;
    case 9: SR_SW(SR_ACT_SOFTPLUS100, 4, false); break;
    case 10: SR_SW(SR_ACT_RELU, 4, false); break;
    case 12: SR_SW(SR_ACT_NONE, 4, true); break;
    case 13: SR_SW(SR_ACT_SOFTPLUS100, 4, true); break;
    case 14: SR_SW(SR_ACT_RELU, 4, true); break;
  }
#undef SR_SW
  if (!kern) return SR_EINVAL;
  static bool attr_set_dev[64][16] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63][key];
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPair);
    if (e1 != cudaSuccess) return (int)e1;
    attr_set = true;
  }
  const int nrp = (MT + 1) / 2;
  const int npairs = nrp < SR_NUM_SMS_B200 / 2 ? nrp : SR_NUM_SMS_B200 / 2;
  kern<<<(unsigned)(2 * npairs), epi_threads(body_mul != 0), kSmemPair, s>>>(sw);
  return sr_launch_status();
}
}
