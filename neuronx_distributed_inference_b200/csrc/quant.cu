// Dynamic per-token fp8 (e4m3) activation quantisation, optionally fused with RMSNorm:
//   y = rmsnorm(x) * gamma (or x),  scale[row] = min(amax(|y|), clamp) / 448,  q = fp8(y / scale[row])
// One CTA per row, the row stays in registers between the three reductions (sum of squares, amax) and the conversion.
// reference kernel: K6 rmsnorm_quant (models/llama/modeling_llama.py:553-575 feeds the fp8 MLP kernels with it).
#include <cuda_fp8.h>

#include <stdexcept>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int RQ_THREADS = 256;
constexpr int RQ_MAXV = 8;   // 16-byte vectors per thread: H <= 8 * 8 * 256 = 16384

__device__ __forceinline__ float block_reduce(float v, float* sred, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  float r = sred[0];
#pragma unroll
  for (int w = 1; w < RQ_THREADS / 32; ++w) r = is_max ? fmaxf(r, sred[w]) : r + sred[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(RQ_THREADS) rmsnorm_quant_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                                                   uint8_t* __restrict__ q, float* __restrict__ scale, int H, float eps,
                                                                   float offset, float clamp) {
  __shared__ float sred[RQ_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const int nvec = H >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * H);
  float v[RQ_MAXV][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < RQ_MAXV; ++j) {
    const int i = tid + j * RQ_THREADS;
    if (i < nvec) {
      const uint4 w = ldg_act(xr + i);
      v[j][0] = bf16lo(w.x); v[j][1] = bf16hi(w.x); v[j][2] = bf16lo(w.y); v[j][3] = bf16hi(w.y);
      v[j][4] = bf16lo(w.z); v[j][5] = bf16hi(w.z); v[j][6] = bf16lo(w.w); v[j][7] = bf16hi(w.w);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[j][e] * v[j][e];
    }
  }
  if (gamma != nullptr) {
    const float rstd = rsqrtf(block_reduce(ss, sred, false) / (float)H + eps);
    const uint4* gr = reinterpret_cast<const uint4*>(gamma);
#pragma unroll
    for (int j = 0; j < RQ_MAXV; ++j) {
      const int i = tid + j * RQ_THREADS;
      if (i < nvec) {
        const uint4 g = ldg_cached(gr + i);
        const float gf[8] = {bf16lo(g.x), bf16hi(g.x), bf16lo(g.y), bf16hi(g.y), bf16lo(g.z), bf16hi(g.z), bf16lo(g.w), bf16hi(g.w)};
#pragma unroll
        for (int e = 0; e < 8; ++e)   // rounded to bf16 like the unfused rmsnorm -> quantise pipeline
          v[j][e] = __bfloat162float(__float2bfloat16(v[j][e] * rstd * (gf[e] + offset)));
      }
    }
  }
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < RQ_MAXV; ++j)
    if (tid + j * RQ_THREADS < nvec)
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(v[j][e]));
  amax = fminf(block_reduce(amax, sred, true), clamp);
  const float sc = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / sc;
  if (tid == 0) scale[row] = sc;
  uint2* qr = reinterpret_cast<uint2*>(q + (size_t)row * H);
#pragma unroll
  for (int j = 0; j < RQ_MAXV; ++j) {
    const int i = tid + j * RQ_THREADS;
    if (i < nvec) {
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float a = fminf(fmaxf(v[j][e] * inv, -448.f), 448.f), b = fminf(fmaxf(v[j][e + 1] * inv, -448.f), 448.f);
        const uint32_t pr = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
        if (e < 4) lo |= pr << (8 * e); else hi |= pr << (8 * (e - 4));
      }
      qr[i] = make_uint2(lo, hi);
    }
  }
}

void rmsnorm_quant_launch(const void* x, const void* gamma, void* q, float* scale, int rows, int H, float eps, float offset, float clamp,
                          cudaStream_t stream) {
  if (H % 8 != 0 || H > RQ_MAXV * 8 * RQ_THREADS) throw std::runtime_error("rmsnorm_quant: hidden must be a multiple of 8 and <= 16384");
  launch_pdl(rmsnorm_quant_kernel, dim3(rows), dim3(RQ_THREADS), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
             reinterpret_cast<const __nv_bfloat16*>(gamma), reinterpret_cast<uint8_t*>(q), scale, H, eps, offset, clamp);
}

}  // namespace nxdi
