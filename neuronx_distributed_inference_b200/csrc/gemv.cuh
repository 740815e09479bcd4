// Weight-streaming skinny GEMM ("tensor-core GEMV") for T <= 8 tokens:  Y[T,N] = f(norm(X)[T,K] · W[N,K]^T)
//
// Decode is HBM-bound on the weights, so the kernel is organised around keeping >=64 KB of 128-bit
// weight loads in flight per SM and doing almost no ALU work per byte:
//   * a CTA (8 warps) owns a 16-row tile of W; warp w streams the k-slice [w*K/8,(w+1)*K/8) of those rows
//     straight from global memory into mma.sync A fragments (no shared-memory staging of W at all: lane
//     (g,t) loads 16 B of row g and 16 B of row g+8 at k-offset 8t, which is a valid A fragment of two
//     m16n8k16 MMAs under a fixed permutation of k that X's B fragment follows);
//   * X (<= 8 tokens, optionally RMS-normalised in the prologue, fused input norm) lives in shared memory
//     as bf16 and supplies the B fragments (token = MMA column), so all tokens ride along for free;
//   * register double-buffering over a *flattened* (tile, k-iteration) index keeps loads for the next tile
//     in flight while the current tile is reduced across the 8 warps and written out;
//   * the first weight batch is issued BEFORE griddepcontrol.wait (PDL): weights never depend on the
//     previous kernel, so HBM streaming continues across kernel boundaries inside the decode CUDA graph;
//   * epilogues: bias, residual, SwiGLU/GeGLU (tile = 8 gate rows + 8 up rows).
// Role since round 2: the fallback for activations too wide for shared memory (T*K*2 > ~200 KB, e.g. 8 tokens x K=14336);
// everything else — and every fused all-reduce — runs on the TMA-pipelined gemv2.cu.
#pragma once
#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int GEMV_WARPS = 8;
constexpr int GEMV_THREADS = GEMV_WARPS * 32;
constexpr int GEMV_U = 4;  // k-iterations (of 32) per register stage

enum GemvAct { ACT_NONE = 0, ACT_SILU_MUL = 1, ACT_GELU_TANH_MUL = 2, ACT_GELU_MUL = 3 };

__device__ __forceinline__ uint32_t f16x2_from_bf16x2(uint32_t v) {
  __half2 h = __floats2half2_rn(bf16lo(v), bf16hi(v));
  return *reinterpret_cast<uint32_t*>(&h);
}

struct GemvCursor {
  int tile;  // global tile index
  int it;    // k-iteration inside my slice
  const __nv_bfloat16 *p0, *p1;
};

template <bool GLU>
__device__ __forceinline__ void gemv_rows(int tile, int g, int N, int& r0, int& r1) {
  if (GLU) {
    r0 = tile * 8 + g;
    r1 = (N >> 1) + tile * 8 + g;
    r0 = min(r0, (N >> 1) - 1);
    r1 = min(r1, N - 1);
  } else {
    r0 = min(tile * 16 + g, N - 1);
    r1 = min(tile * 16 + g + 8, N - 1);
  }
}

template <bool GLU>
__global__ void __launch_bounds__(GEMV_THREADS, 2) gemv_kernel(const GemvParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int K = p.K, N = p.N, T = p.T;
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(p.w);
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* BIAS = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  const __nv_bfloat16* RES = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);
  const int xs_stride = K * 2 + 64;  // bytes; +64 keeps the 8-lane LDS.128 phases conflict free
  uint8_t* xs = smem_raw;
  float* red = reinterpret_cast<float*>(smem_raw + (p.x_in_smem ? (size_t)T * xs_stride : 0));  // [2][8][128]
  float* rstd_s = red + 2 * GEMV_WARPS * 128;                                                       // [8][8]

  const int kslice = K / GEMV_WARPS;
  const int kbeg = warp * kslice + t4 * 8;
  const int n_iter = kslice >> 5;
  const int n_tiles = GLU ? ((N >> 1) + 7) >> 3 : (N + 15) >> 4;

  // ---- issue cursor -------------------------------------------------------------------------------
  GemvCursor ld;
  ld.tile = blockIdx.x;
  ld.it = 0;
  auto set_ptrs = [&](GemvCursor& c) {
    int r0, r1;
    gemv_rows<GLU>(min(c.tile, n_tiles - 1), g, N, r0, r1);
    c.p0 = W + (size_t)r0 * K + kbeg;
    c.p1 = W + (size_t)r1 * K + kbeg;
  };
  set_ptrs(ld);
  uint4 A0[2][GEMV_U], A1[2][GEMV_U];
  auto issue = [&](int s) {
#pragma unroll
    for (int u = 0; u < GEMV_U; ++u) {
      if (ld.tile < n_tiles) {
        A0[s][u] = ldg_stream(ld.p0 + ld.it * 32);
        A1[s][u] = ldg_stream(ld.p1 + ld.it * 32);
        if (++ld.it == n_iter) {
          ld.it = 0;
          ld.tile += gridDim.x;
          set_ptrs(ld);
        }
      }
    }
  };
  issue(0);  // weights do not depend on the previous kernel: start streaming before the PDL wait
  pdl_launch_dependents();
  pdl_wait();

  // ---- X prologue: (optional RMSNorm) -> bf16 in shared memory -------------------------------------
  if (p.x_in_smem) {
    float ss[GEMV_MAX_T];
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) ss[t] = 0.f;
    const int nvec = K >> 3;
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      if (t >= T) break;
      const uint4* src = reinterpret_cast<const uint4*>(X + (size_t)t * p.ldx);
      uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
      float acc = 0.f;
      for (int v = tid; v < nvec; v += GEMV_THREADS) {
        uint4 q = ldg_cached(src + v);
        dst[v] = q;
        acc += bf16lo(q.x) * bf16lo(q.x) + bf16hi(q.x) * bf16hi(q.x) + bf16lo(q.y) * bf16lo(q.y) +
               bf16hi(q.y) * bf16hi(q.y) + bf16lo(q.z) * bf16lo(q.z) + bf16hi(q.z) * bf16hi(q.z) +
               bf16lo(q.w) * bf16lo(q.w) + bf16hi(q.w) * bf16hi(q.w);
      }
      ss[t] = acc;
    }
    if (p.norm_w != nullptr) {
#pragma unroll
      for (int t = 0; t < GEMV_MAX_T; ++t) {
        if (t < T) {
          float v = warp_sum(ss[t]);
          if (lane == 0) rstd_s[warp * 8 + t] = v;
        }
      }
      __syncthreads();
      for (int t = 0; t < T; ++t) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < GEMV_WARPS; ++w) tot += rstd_s[w * 8 + t];
        const float rstd = rsqrtf(tot / (float)K + p.eps);
        uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
        const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
        for (int v = tid; v < nvec; v += GEMV_THREADS) {
          uint4 q = dst[v];
          uint4 gm = ldg_cached(gw + v);
          const float o = p.norm_offset;
          q.x = pack_bf16(bf16lo(q.x) * rstd * (bf16lo(gm.x) + o), bf16hi(q.x) * rstd * (bf16hi(gm.x) + o));
          q.y = pack_bf16(bf16lo(q.y) * rstd * (bf16lo(gm.y) + o), bf16hi(q.y) * rstd * (bf16hi(gm.y) + o));
          q.z = pack_bf16(bf16lo(q.z) * rstd * (bf16lo(gm.z) + o), bf16hi(q.z) * rstd * (bf16hi(gm.z) + o));
          q.w = pack_bf16(bf16lo(q.w) * rstd * (bf16lo(gm.w) + o), bf16hi(q.w) * rstd * (bf16hi(gm.w) + o));
          dst[v] = q;
        }
      }
    }
    __syncthreads();
  }

  // ---- main loop over the flattened (tile, k-iteration) space --------------------------------------
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
  int ctile = blockIdx.x, cit = 0, parity = 0;
  const bool tok_ok = g < T;
  const uint8_t* xrow = p.x_in_smem ? xs + (size_t)g * xs_stride
                                    : reinterpret_cast<const uint8_t*>(X + (size_t)min(g, T - 1) * p.ldx);

  auto flush = [&]() {
    // cross-warp reduction of the 16x8 tile, then epilogue by 128 threads (row fastest)
    float* r = red + parity * (GEMV_WARPS * 128) + warp * 128;
    r[g * 8 + 2 * t4] = c0[0] + c1[0];
    r[g * 8 + 2 * t4 + 1] = c0[1] + c1[1];
    r[(g + 8) * 8 + 2 * t4] = c0[2] + c1[2];
    r[(g + 8) * 8 + 2 * t4 + 1] = c0[3] + c1[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) c0[i] = c1[i] = 0.f;
    __syncthreads();
    const float* rb = red + parity * (GEMV_WARPS * 128);
    if (tid < 128) {
      const int col = tid >> 4, row = tid & 15;
      if (col < T) {
        if (GLU) {
          if (row < 8) {
            float gate = 0.f, up = 0.f;
#pragma unroll
            for (int w = 0; w < GEMV_WARPS; ++w) {
              gate += rb[w * 128 + row * 8 + col];
              up += rb[w * 128 + (row + 8) * 8 + col];
            }
            const int n = ctile * 8 + row, half = N >> 1;
            if (n < half) {
              if (p.bias != nullptr) {
                gate += __bfloat162float(BIAS[n]);
                up += __bfloat162float(BIAS[half + n]);
              }
              float a = p.act == ACT_SILU_MUL ? silu(gate) : (p.act == ACT_GELU_TANH_MUL ? gelu_tanh(gate) : gelu_erf(gate));
              Y[(size_t)col * p.ldy + n] = __float2bfloat16(a * up);
            }
          }
        } else {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < GEMV_WARPS; ++w) v += rb[w * 128 + row * 8 + col];
          const int n = ctile * 16 + row;
          if (n < N) {
            if (p.bias != nullptr) v += __bfloat162float(BIAS[n]);
            if (p.residual != nullptr) v += __bfloat162float(RES[(size_t)col * p.ldy + n]);
            Y[(size_t)col * p.ldy + n] = __float2bfloat16(v);
          }
        }
      }
    }
    parity ^= 1;
  };

  auto consume = [&](int s) {
#pragma unroll
    for (int u = 0; u < GEMV_U; ++u) {
      if (ctile < n_tiles) {
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (tok_ok) {
          const uint8_t* xp = xrow + (size_t)(kbeg + cit * 32) * 2;
          xv = p.x_in_smem ? *reinterpret_cast<const uint4*>(xp) : ldg_cached(xp);
        }
        const uint4 a0 = A0[s][u], a1 = A1[s][u];
        {
          const uint32_t a[4] = {a0.x, a1.x, a0.y, a1.y};
          const uint32_t b[2] = {xv.x, xv.y};
          mma_bf16_16816(c0, a, b);
        }
        {
          const uint32_t a[4] = {a0.z, a1.z, a0.w, a1.w};
          const uint32_t b[2] = {xv.z, xv.w};
          mma_bf16_16816(c1, a, b);
        }
        if (++cit == n_iter) {
          flush();
          cit = 0;
          ctile += gridDim.x;
        }
      }
    }
  };

  while (ctile < n_tiles) {
    issue(1);
    consume(0);
    issue(0);
    consume(1);
  }
}


}  // namespace nxdi
