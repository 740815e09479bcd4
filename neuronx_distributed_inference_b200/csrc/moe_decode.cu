// MoE token-generation kernels: the routed experts of a decode step (T <= 8 tokens, top-k experts each) as two batched
// weight-streaming launches over (token, slot) pairs — no host-side dispatch, no data-dependent shapes, CUDA-graph safe.
//   K1  u[p, :] = silu(x_t . Wg_e^T) * (x_t . Wu_e^T)        W_gate_up [E, 2I, H]  (rows [0,I) gate, [I,2I) up)
//   K2  y[p, :] = u[p, :] . Wd_e^T                           W_down    [E, H, I]
// with p = (t, j), e = topk_i[t, j] - expert_offset; pairs routed to an expert that lives on another EP rank produce zeros.
// The combine  out[t] = sum_j w[t, j] * y[(t, j)]  is fused into K2's epilogue through fp32 atomics on a zeroed [T, H] buffer.
// reference kernels replaced: K16 moe_token_gen_all_experts / fused MoE TKG kernels (moe_v2.py), K17 for T <= 8.
//
// Design: one warp per output row, lanes stride the reduction dimension with 16-byte loads (x / u row staged in shared
// memory once per CTA); 4-way unrolled so every lane keeps 8 independent 16-byte weight loads in flight.  grid = (row
// blocks, pairs): when two tokens pick the same expert the second read of its weights comes from L2.
#include <stdexcept>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int MOE_THREADS = 256;

template <bool GLU>
__global__ void __launch_bounds__(MOE_THREADS) moe_rows_kernel(const __nv_bfloat16* __restrict__ xin,   // K1: x [T,K]; K2: u [P,K]
                                                                const __nv_bfloat16* __restrict__ W,     // [E, N_rows, K]
                                                                const int* __restrict__ topk_i,           // [P]
                                                                const float* __restrict__ topk_w,         // [P] (K2 only)
                                                                __nv_bfloat16* __restrict__ u_out,        // K1: [P, N]
                                                                float* __restrict__ y_acc,                // K2: [T, N] fp32 (atomics)
                                                                int K, int N, int topk, int E, int expert_offset) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* xs = reinterpret_cast<uint4*>(smem);
  pdl_launch_dependents();
  pdl_wait();
  const int p = blockIdx.y, t = p / topk;
  const int e = topk_i[p] - expert_offset;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows_per = (N + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(N, r0 + rows_per);
  if (e < 0 || e >= E) {   // expert owned by another rank
    if (GLU)
      for (int r = r0 + tid; r < r1; r += MOE_THREADS) u_out[(size_t)p * N + r] = __float2bfloat16(0.f);
    return;
  }
  const int nvec = K >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(xin + (size_t)(GLU ? t : p) * K);
  for (int v = tid; v < nvec; v += MOE_THREADS) xs[v] = __ldcg(src + v);
  __syncthreads();
  const size_t rows_total = GLU ? (size_t)2 * N : (size_t)N;
  const __nv_bfloat16* We = W + (size_t)e * rows_total * K;
  auto dot8 = [](const uint4& a, const uint4& b) {
    return bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y) +
           bf16lo(a.z) * bf16lo(b.z) + bf16hi(a.z) * bf16hi(b.z) + bf16lo(a.w) * bf16lo(b.w) + bf16hi(a.w) * bf16hi(b.w);
  };
  for (int r = r0 + warp; r < r1; r += MOE_THREADS / 32) {
    const uint4* wa = reinterpret_cast<const uint4*>(We + (size_t)r * K);
    const uint4* wb = GLU ? reinterpret_cast<const uint4*>(We + (size_t)(N + r) * K) : nullptr;
    float acc_a = 0.f, acc_b = 0.f;
    int v = lane;
    for (; v + 96 < nvec; v += 128) {   // 4 independent 16-byte loads per operand in flight
      const uint4 a0 = ldg_stream(wa + v), a1 = ldg_stream(wa + v + 32), a2 = ldg_stream(wa + v + 64), a3 = ldg_stream(wa + v + 96);
      uint4 b0, b1, b2, b3;
      if (GLU) { b0 = ldg_stream(wb + v); b1 = ldg_stream(wb + v + 32); b2 = ldg_stream(wb + v + 64); b3 = ldg_stream(wb + v + 96); }
      const uint4 x0 = xs[v], x1 = xs[v + 32], x2 = xs[v + 64], x3 = xs[v + 96];
      acc_a += dot8(a0, x0) + dot8(a1, x1) + dot8(a2, x2) + dot8(a3, x3);
      if (GLU) acc_b += dot8(b0, x0) + dot8(b1, x1) + dot8(b2, x2) + dot8(b3, x3);
    }
    for (; v < nvec; v += 32) {
      const uint4 xv = xs[v];
      acc_a += dot8(ldg_stream(wa + v), xv);
      if (GLU) acc_b += dot8(ldg_stream(wb + v), xv);
    }
    acc_a = warp_sum(acc_a);
    if (GLU) acc_b = warp_sum(acc_b);
    if (lane == 0) {
      if (GLU) u_out[(size_t)p * N + r] = __float2bfloat16(silu(acc_a) * acc_b);
      else atomicAdd(y_acc + (size_t)t * N + r, topk_w[p] * acc_a);
    }
  }
}

void moe_decode_launch(const void* x, const void* w_gate_up, const void* w_down, const float* topk_w, const int* topk_i, void* u,
                       float* y_acc, int T, int topk, int H, int I, int E, int expert_offset, int n_sms, cudaStream_t stream) {
  const int P = T * topk;
  if (H % 8 != 0 || I % 8 != 0) throw std::runtime_error("moe_decode: H and I must be multiples of 8");
  const int gx = std::max(1, (4 * n_sms + P - 1) / P);
  auto k1 = moe_rows_kernel<true>;
  auto k2 = moe_rows_kernel<false>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured = true;
  }
  if ((size_t)H * 2 > 64 * 1024 || (size_t)I * 2 > 64 * 1024) throw std::runtime_error("moe_decode: row too wide for shared memory");
  launch_pdl(k1, dim3(std::min(gx, I), P), dim3(MOE_THREADS), (size_t)H * 2, stream, reinterpret_cast<const __nv_bfloat16*>(x),
             reinterpret_cast<const __nv_bfloat16*>(w_gate_up), topk_i, topk_w, reinterpret_cast<__nv_bfloat16*>(u),
             static_cast<float*>(nullptr), H, I, topk, E, expert_offset);
  launch_pdl(k2, dim3(std::min(gx, H), P), dim3(MOE_THREADS), (size_t)I * 2, stream, reinterpret_cast<const __nv_bfloat16*>(u),
             reinterpret_cast<const __nv_bfloat16*>(w_down), topk_i, topk_w, static_cast<__nv_bfloat16*>(nullptr), y_acc, I, H, topk,
             E, expert_offset);
}

}  // namespace nxdi
