// Shared pieces of the tensor-core engine (tc_gemm.cu: layer GEMMs, tc_wgrad.cu: weight-gradient GEMMs):
// tile geometry of the pre-tiled split-bf16 operands, tcgen05 / TMEM wrappers.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace sr_tc {

#ifndef SR_TC_PLANES
#define SR_TC_PLANES 2
#endif
constexpr int kPlanes = SR_TC_PLANES;
constexpr int BM = 128, BN = 256, BK = 32, STAGES = kPlanes == 2 ? 4 : 3;
// Epilogue warps per CTA, per kernel family: the forward epilogues (bias, activation, re-split) are latency-bound with
// two warps per scheduler and fit 96 registers, so they run 16 warps (four per TMEM lane quarter, 64 columns each:
// measured -7 % per layer at M = 50 333); the reverse epilogues prefetch their operand tiles and need the 168-register
// budget of 8 warps (at 16 they spill: +40 %).
#ifndef SR_TC_EPI_WARPS_FWD
#define SR_TC_EPI_WARPS_FWD 16
#endif
#ifndef SR_TC_EPI_WARPS_REV
#define SR_TC_EPI_WARPS_REV 8
#endif
__host__ __device__ constexpr int epi_warps(bool mul) { return mul ? SR_TC_EPI_WARPS_REV : SR_TC_EPI_WARPS_FWD; }
__host__ __device__ constexpr int epi_threads(bool mul) { return 64 + 32 * epi_warps(mul); }
__host__ __device__ constexpr int epi_part_cols(int ew) { return 256 / (ew / 4); }     // accumulator columns per epilogue warp
__host__ __device__ constexpr int epi_chunks(int ew) { return epi_part_cols(ew) / 32; }  // 32-column chunks per warp
static_assert((SR_TC_EPI_WARPS_FWD == 8 || SR_TC_EPI_WARPS_FWD == 16) &&
              (SR_TC_EPI_WARPS_REV == 8 || SR_TC_EPI_WARPS_REV == 16), "epilogue warps: 8 or 16");
constexpr int kMaxEpiWarps = 16;
constexpr int A_PLANE = BM * BK;          // elements
constexpr int W_PLANE = BN * BK;
constexpr int A_STAGE = kPlanes * A_PLANE;   // 2 planes: 16 KB
constexpr int W_STAGE = kPlanes * W_PLANE;   // 2 planes: 32 KB
constexpr uint32_t A_STAGE_BYTES = A_STAGE * 2, W_STAGE_BYTES = W_STAGE * 2;
constexpr size_t kSmem = (size_t)STAGES * (A_STAGE_BYTES + W_STAGE_BYTES) + 256;

// ---- tiled ("pre-swizzled") global layouts ----------------------------------------------------
// A: [row tile mt][k chunk kc][plane p][k8 (4)][row group (16)][row (8)][elem (8)]
// W: [col tile nt][k chunk kc][plane p][k8 (4)][row group (32)][row (8)][elem (8)]
__host__ __device__ inline size_t a_tile_off(long long mt, int kc, int KC, int p) {
  return (((size_t)mt * KC + kc) * kPlanes + p) * A_PLANE;
}
__host__ __device__ inline size_t w_tile_off(int nt, int kc, int KC, int p) {
  return (((size_t)nt * KC + kc) * kPlanes + p) * W_PLANE;
}
__device__ __forceinline__ int in_tile_off(int rows_per_tile, int r, int k) {
  return (k >> 3) * (rows_per_tile * 8) + (r >> 3) * 64 + (r & 7) * 8 + (k & 7);
}

__device__ __forceinline__ void split3(float x, __nv_bfloat16& b1, __nv_bfloat16& b2, __nv_bfloat16& b3) {
  b1 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(b1);
  b2 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(b2);
  b3 = __float2bfloat16_rn(r2);
}

// ---- tcgen05 wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1, swizzle none
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t kIdescBase = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   sr_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#define SR_TMEM_REGS32(v)                                                                          \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),  \
  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
  "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
  "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
  "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define SR_TMEM_REGS32_RW(v)                                                                       \
  "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),  \
  "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),          \
  "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),        \
  "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),        \
  "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
// asynchronous TMEM -> register load of 32 columns (this warp's 32 lanes); pair with tmem_wait
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : SR_TMEM_REGS32(v)
      : "r"(taddr));
}
// the registers are in/out operands so that no consumer can be scheduled above the wait
__device__ __forceinline__ void tmem_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : SR_TMEM_REGS32_RW(v)::"memory");
}

}  // namespace sr_tc
