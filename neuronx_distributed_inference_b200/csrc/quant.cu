// Dynamic per-token fp8 (e4m3) activation quantisation, optionally fused with RMSNorm:
//   y = rmsnorm(x) * gamma (or x),  scale[row] = min(amax(|y|), clamp) / 448,  q = fp8(y / scale[row])
// One CTA per row, the row stays in registers between the three reductions (sum of squares, amax) and the conversion.
// reference kernel: K6 rmsnorm_quant (models/llama/modeling_llama.py:553-575 feeds the fp8 MLP kernels with it).
#include <cuda_fp8.h>

#include <algorithm>
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

// Weight-only 8-bit layers at prefill sizes without activation quantisation: expand W (int8 / fp8-e4m3, per-channel or per-tensor
// scale) to bf16 ONCE per call into a scratch buffer (1 B read + 2 B written per weight; the copy is consumed from L2 by the
// tcgen05 GEMM that follows) instead of the three elementwise passes of the PyTorch composite.
template <int WT>
__global__ void __launch_bounds__(256) dequant_bf16_kernel(const uint8_t* __restrict__ w, const float* __restrict__ scale, int scale_n,
                                                           __nv_bfloat16* __restrict__ out, int N, int K) {
  pdl_launch_dependents();
  const int kv = K >> 4;                       // 16-byte vectors per row
  const long long total = (long long)N * kv;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int n = (int)(i / kv);
    const float sc = scale[scale_n == 1 ? 0 : n];
    const uint4 q = ldg_stream(reinterpret_cast<const uint4*>(w) + i);
    const uint32_t words[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[4];
      if (WT == 1) {
        f[0] = (float)(int8_t)(words[j] & 0xff); f[1] = (float)(int8_t)((words[j] >> 8) & 0xff);
        f[2] = (float)(int8_t)((words[j] >> 16) & 0xff); f[3] = (float)(int8_t)(words[j] >> 24);
      } else {
        const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(words[j] & 0xffff), __NV_E4M3);
        const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(words[j] >> 16), __NV_E4M3);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo)), b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
      }
      o[2 * j] = pack_bf16(f[0] * sc, f[1] * sc);
      o[2 * j + 1] = pack_bf16(f[2] * sc, f[3] * sc);
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + 2 * i;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

void dequant_bf16_launch(const void* w, int wdtype, const float* scale, int scale_n, void* out, int N, int K, cudaStream_t stream) {
  if (K % 16 != 0) throw std::runtime_error("dequant_bf16: K must be a multiple of 16");
  const long long total = (long long)N * (K / 16);
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (wdtype == 1)
    launch_pdl(dequant_bf16_kernel<1>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint8_t*>(w), scale, scale_n,
               reinterpret_cast<__nv_bfloat16*>(out), N, K);
  else
    launch_pdl(dequant_bf16_kernel<2>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint8_t*>(w), scale, scale_n,
               reinterpret_cast<__nv_bfloat16*>(out), N, K);
}

}  // namespace nxdi
