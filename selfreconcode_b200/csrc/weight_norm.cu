// Effective weights of weight-normalised linear layers, all layers of a network in ONE launch, with backward.
//   w[r, :] = g[r] * v[r, :] / ||v[r, :]||            (torch.nn.utils.weight_norm, dim = 0; model/network.py:60-61)
//   backward:  s = <gw[r], v[r]>;  gg[r] = s / ||v||;  gv[r, :] = (g / ||v||) * (gw[r, :] - v[r, :] * s / ||v||^2)
// One CTA per row (rows are 39..512 floats wide): HBM stream, 8 B/element forward, 16 B/element backward.  Replaces ~12
// element-wise / reduce launches per layer per direction in the optimisation step.
#include "common.cuh"

namespace {

constexpr int kWnMaxLayers = 12;
struct WnArgs {
  sr_wn_layer layer[kWnMaxLayers];
  int row0[kWnMaxLayers + 1];   // prefix sums of the row counts: blockIdx.x -> (layer, row)
  int L;
};

__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();                 // `red` may still be read from a previous call
  if (l == 0) red[w] = x;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) t += red[i];     // 128 threads: fixed order, deterministic
  return t;
}

__device__ __forceinline__ bool locate(const WnArgs& a, int& li, int& row) {
  const int b = blockIdx.x;
  li = 0;
  while (li < a.L && b >= a.row0[li + 1]) ++li;
  if (li >= a.L) return false;
  row = b - a.row0[li];
  return true;
}

__global__ void __launch_bounds__(128) wn_fwd_kernel(const __grid_constant__ WnArgs a) {
  __shared__ float red[4];
  int li, row;
  if (!locate(a, li, row)) return;
  const sr_wn_layer& L = a.layer[li];
  const float* v = L.v + (size_t)row * L.k;
  float ss = 0.f;
  for (int c = threadIdx.x; c < L.k; c += 128) { const float x = v[c]; ss += x * x; }
  ss = block_sum(ss, red);
  const float nrm = sqrtf(ss);                 // ||v||
  const float sc = L.g[row] / nrm;
  float* w = L.w + (size_t)row * L.k;
  for (int c = threadIdx.x; c < L.k; c += 128) w[c] = v[c] * sc;
  if (threadIdx.x == 0) L.inv_norm[row] = 1.0f / nrm;
}

__global__ void __launch_bounds__(128) wn_bwd_kernel(const __grid_constant__ WnArgs a) {
  __shared__ float red[4];
  int li, row;
  if (!locate(a, li, row)) return;
  const sr_wn_layer& L = a.layer[li];
  const float* v = L.v + (size_t)row * L.k;
  float* gv = L.gv + (size_t)row * L.k;
  if (L.gw == nullptr) {                        // this layer's weights received no gradient
    for (int c = threadIdx.x; c < L.k; c += 128) gv[c] = 0.f;
    if (threadIdx.x == 0) L.gg[row] = 0.f;
    return;
  }
  const float* gw = L.gw + (size_t)row * L.gw_ld;
  float s = 0.f;
  for (int c = threadIdx.x; c < L.k; c += 128) s += gw[c] * v[c];
  s = block_sum(s, red);
  const float inv = L.inv_norm[row];
  const float gi = L.g[row] * inv, t = s * inv * inv;
  for (int c = threadIdx.x; c < L.k; c += 128) gv[c] = gi * (gw[c] - v[c] * t);
  if (threadIdx.x == 0) L.gg[row] = s * inv;
}

int fill(WnArgs& a, const sr_wn_layer* layers, int L, bool bwd) {
  if (!layers || L <= 0 || L > kWnMaxLayers) return SR_EINVAL;
  a.L = L;
  a.row0[0] = 0;
  for (int i = 0; i < L; ++i) {
    const sr_wn_layer& l = layers[i];
    if (!l.v || !l.g || !l.inv_norm || l.n <= 0 || l.k <= 0) return SR_EINVAL;
    if (!bwd && !l.w) return SR_EINVAL;
    if (bwd && (!l.gv || !l.gg || (l.gw && l.gw_ld < l.k))) return SR_EINVAL;
    a.layer[i] = l;
    a.row0[i + 1] = a.row0[i] + l.n;
  }
  return SR_OK;
}

}  // namespace

extern "C" {

int sr_weight_norm_forward(const sr_wn_layer* layers, int L, cudaStream_t s) {
  WnArgs a;
  const int e = fill(a, layers, L, false);
  if (e != SR_OK) return e;
  wn_fwd_kernel<<<a.row0[L], 128, 0, s>>>(a);
  return sr_launch_status();
}

int sr_weight_norm_backward(const sr_wn_layer* layers, int L, cudaStream_t s) {
  WnArgs a;
  const int e = fill(a, layers, L, true);
  if (e != SR_OK) return e;
  wn_bwd_kernel<<<a.row0[L], 128, 0, s>>>(a);
  return sr_launch_status();
}
}
