// Weight-only quantised skinny GEMM for decode:  Y[T<=8, N] = epi( norm(X)[T,K] . dequant(Wq[N,K])^T )
//   Wq: int8 or fp8-e4m3 (one byte per weight), scale: fp32 per output channel ([N]) or per tensor ([1]).
// Decode is bound by weight bytes; 8-bit weights halve them, but only if nothing materialises a bf16 copy — the torch
// fallback dequantises the whole matrix (1 + 2 + 2 bytes of traffic per weight).  Here the bytes go HBM -> registers once:
//   * warp-per-row-group: a warp owns R consecutive output rows and accumulates T x R fp32 sums; lanes stride K with
//     16-byte loads (16 weights), four loads in flight per lane and row;
//   * X (optionally multiplied by the RMSNorm gamma; 1/rms is applied in the epilogue) is staged in shared memory in K
//     chunks of 4096 columns, so any K fits;
//   * epilogue: x scale[n], + bias, SwiGLU / GeGLU over (gate row n, up row N/2 + n), + residual, bf16 store.
// reference kernels replaced: the quantised flavours of K3/K4/K5 (decode), K6 rmsnorm_quant's consumer.
#include <cuda_fp8.h>

#include <stdexcept>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int QG_THREADS = 256, QG_WARPS = 8, QG_KC = 4096;

template <int WT>  // 1 int8, 2 fp8 e4m3
__device__ __forceinline__ void dequant16(const uint4& q, float (&f)[16]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (WT == 1) {
      f[4 * i + 0] = (float)(int8_t)(w[i] & 0xff);
      f[4 * i + 1] = (float)(int8_t)((w[i] >> 8) & 0xff);
      f[4 * i + 2] = (float)(int8_t)((w[i] >> 16) & 0xff);
      f[4 * i + 3] = (float)(int8_t)((w[i] >> 24) & 0xff);
    } else {
      const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] & 0xffff), __NV_E4M3);
      const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] >> 16), __NV_E4M3);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
      f[4 * i + 0] = a.x; f[4 * i + 1] = a.y; f[4 * i + 2] = b.x; f[4 * i + 3] = b.y;
    }
  }
}

struct QGemvParams {
  const __nv_bfloat16* x;
  const uint8_t* w;
  const float* scale;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* norm_w;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  int T, N, K, ldx, ldy, act, scale_n, rows_per_warp;
  float eps, norm_offset;
};

template <int WT, int R, bool GLU>
__global__ void __launch_bounds__(QG_THREADS) qgemv_kernel(const QGemvParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);                  // [T][QG_KC]
  __shared__ float ss_part[QG_WARPS][GEMV_MAX_T];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, K = p.K;
  const int n_log = GLU ? (p.N >> 1) : p.N;                                    // logical output rows
  const int row0 = (blockIdx.x * QG_WARPS + warp) * R;
  pdl_launch_dependents();
  pdl_wait();
  float acc[GLU ? 2 : 1][R][GEMV_MAX_T];
#pragma unroll
  for (int h = 0; h < (GLU ? 2 : 1); ++h)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int t = 0; t < GEMV_MAX_T; ++t) acc[h][r][t] = 0.f;
  float ss[GEMV_MAX_T];
#pragma unroll
  for (int t = 0; t < GEMV_MAX_T; ++t) ss[t] = 0.f;

  for (int k0 = 0; k0 < K; k0 += QG_KC) {
    const int kc = min(QG_KC, K - k0);
    __syncthreads();
    // ---- stage x[:, k0:k0+kc] (x gamma) ----
    for (int t = 0; t < T; ++t) {
      const uint4* src = reinterpret_cast<const uint4*>(p.x + (size_t)t * p.ldx + k0);
      uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * QG_KC);
      for (int v = tid; v < (kc >> 3); v += QG_THREADS) {
        uint4 q = ldg_cached(src + v);
        if (p.norm_w != nullptr) {
          ss[t] += bf16lo(q.x) * bf16lo(q.x) + bf16hi(q.x) * bf16hi(q.x) + bf16lo(q.y) * bf16lo(q.y) + bf16hi(q.y) * bf16hi(q.y) +
                   bf16lo(q.z) * bf16lo(q.z) + bf16hi(q.z) * bf16hi(q.z) + bf16lo(q.w) * bf16lo(q.w) + bf16hi(q.w) * bf16hi(q.w);
          const uint4 gm = ldg_cached(reinterpret_cast<const uint4*>(p.norm_w + k0) + v);
          const float o = p.norm_offset;
          q.x = pack_bf16(bf16lo(q.x) * (bf16lo(gm.x) + o), bf16hi(q.x) * (bf16hi(gm.x) + o));
          q.y = pack_bf16(bf16lo(q.y) * (bf16lo(gm.y) + o), bf16hi(q.y) * (bf16hi(gm.y) + o));
          q.z = pack_bf16(bf16lo(q.z) * (bf16lo(gm.z) + o), bf16hi(q.z) * (bf16hi(gm.z) + o));
          q.w = pack_bf16(bf16lo(q.w) * (bf16lo(gm.w) + o), bf16hi(q.w) * (bf16hi(gm.w) + o));
        }
        dst[v] = q;
      }
    }
    __syncthreads();
    // ---- accumulate: 16 weights per lane and step ----
    const int nv16 = kc >> 4;
#pragma unroll
    for (int h = 0; h < (GLU ? 2 : 1); ++h) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = row0 + r;
        if (row >= n_log) continue;
        const uint4* wr = reinterpret_cast<const uint4*>(p.w + (size_t)(h * n_log + row) * K + k0);
        for (int v = lane; v < nv16; v += 32) {
          float f[16];
          dequant16<WT>(ldg_stream(wr + v), f);
#pragma unroll
          for (int t = 0; t < GEMV_MAX_T; ++t) {
            if (t >= T) break;
            const uint4* xv = reinterpret_cast<const uint4*>(xs + (size_t)t * QG_KC + v * 16);
            const uint4 a = xv[0], b = xv[1];
            acc[h][r][t] += f[0] * bf16lo(a.x) + f[1] * bf16hi(a.x) + f[2] * bf16lo(a.y) + f[3] * bf16hi(a.y) +
                            f[4] * bf16lo(a.z) + f[5] * bf16hi(a.z) + f[6] * bf16lo(a.w) + f[7] * bf16hi(a.w) +
                            f[8] * bf16lo(b.x) + f[9] * bf16hi(b.x) + f[10] * bf16lo(b.y) + f[11] * bf16hi(b.y) +
                            f[12] * bf16lo(b.z) + f[13] * bf16hi(b.z) + f[14] * bf16lo(b.w) + f[15] * bf16hi(b.w);
          }
        }
      }
    }
  }
  // ---- 1/rms per token ----
  float rstd[GEMV_MAX_T];
#pragma unroll
  for (int t = 0; t < GEMV_MAX_T; ++t) rstd[t] = 1.f;
  if (p.norm_w != nullptr) {
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      const float v = warp_sum(ss[t]);
      if (lane == 0) ss_part[warp][t] = v;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < QG_WARPS; ++w) tot += ss_part[w][t];
      rstd[t] = rsqrtf(tot / (float)K + p.eps);
    }
  }
  // ---- epilogue ----
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row0 + r;
    if (row >= n_log) continue;
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      if (t >= T) break;
      float a = warp_sum(acc[0][r][t]);
      float b = GLU ? warp_sum(acc[1][r][t]) : 0.f;
      if (lane != 0) continue;
      const float s0 = p.scale[p.scale_n == 1 ? 0 : row];
      a *= s0 * rstd[t];
      if (p.bias != nullptr) a += __bfloat162float(p.bias[row]);
      float out;
      if (GLU) {
        const float s1 = p.scale[p.scale_n == 1 ? 0 : n_log + row];
        b *= s1 * rstd[t];
        if (p.bias != nullptr) b += __bfloat162float(p.bias[n_log + row]);
        const float g = p.act == 1 ? silu(a) : (p.act == 2 ? gelu_tanh(a) : gelu_erf(a));
        out = g * b;
      } else {
        out = a;
        if (p.residual != nullptr) out += __bfloat162float(p.residual[(size_t)t * p.ldy + row]);
      }
      p.y[(size_t)t * p.ldy + row] = __float2bfloat16(out);
    }
  }
}

template <int WT, int R, bool GLU>
static void launch_q(const QGemvParams& p, int n_log, cudaStream_t stream) {
  auto kern = qgemv_kernel<WT, R, GLU>;
  const size_t smem = (size_t)p.T * QG_KC * 2;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMV_MAX_T * QG_KC * 2);
    configured = true;
  }
  const int grid = (n_log + QG_WARPS * R - 1) / (QG_WARPS * R);
  launch_pdl(kern, dim3(grid), dim3(QG_THREADS), smem, stream, p);
}

void qgemv_launch(const void* x, const void* w, const float* scale, int scale_n, const void* bias, const void* norm_w,
                  const void* residual, void* y, int T, int N, int K, int ldx, int ldy, int act, int wdtype, float eps,
                  float norm_offset, int n_sms, cudaStream_t stream) {
  if (K % 16 != 0) throw std::runtime_error("qgemv: K must be a multiple of 16");
  QGemvParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.w = reinterpret_cast<const uint8_t*>(w);
  p.scale = scale;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.norm_w = reinterpret_cast<const __nv_bfloat16*>(norm_w);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.T = T; p.N = N; p.K = K; p.ldx = ldx; p.ldy = ldy; p.act = act; p.scale_n = scale_n; p.eps = eps; p.norm_offset = norm_offset;
  const bool glu = act != 0;
  const int n_log = glu ? N / 2 : N;
  // rows per warp: keep >= ~4 CTAs per SM worth of parallelism, up to 2 rows per warp for very tall matrices
  const bool two = (long long)n_log >= (long long)n_sms * 4 * QG_WARPS * 2 && T <= 4;
  p.rows_per_warp = two ? 2 : 1;
#define QG_DISPATCH(WT)                                                        \
  if (glu) { if (two) launch_q<WT, 2, true>(p, n_log, stream); else launch_q<WT, 1, true>(p, n_log, stream); } \
  else { if (two) launch_q<WT, 2, false>(p, n_log, stream); else launch_q<WT, 1, false>(p, n_log, stream); }
  if (wdtype == 1) { QG_DISPATCH(1) }
  else if (wdtype == 2) { QG_DISPATCH(2) }
  else throw std::runtime_error("qgemv: weights must be int8 or fp8-e4m3");
#undef QG_DISPATCH
}

}  // namespace nxdi
