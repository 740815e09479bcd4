// RoPE + (optional per-head q/k RMSNorm) + KV-cache append in one pass over the fused QKV projection
// output, plus stand-alone contiguous / paged cache appends.
// reference kernels: K10 write_kv_cache_at_batch_kernel (kvcache/utils.py:17-67, OOB-skip semantics),
// the RoPE/QK-norm stages of K2/K3 (attention_base.py:1186-1381, gqa.py:566-631).
#include "api.h"
#include "common.cuh"

namespace nxdi {

// grid: (B*T, n_q + 2*n_kv); block: D/2 threads... one thread handles the rotation pair (i, i + D/2).
template <int D>
__global__ void rope_kv_append_kernel(const __nv_bfloat16* __restrict__ qkv,  // [B*T, (nq+2nkv)*D]
                                      const float* __restrict__ cos, const float* __restrict__ sin,  // [B*T, D/2]
                                      __nv_bfloat16* __restrict__ q_out,                              // [B*T, nq, D]
                                      __nv_bfloat16* __restrict__ k_out, __nv_bfloat16* __restrict__ v_out,   // [B*T, nkv, D] or null
                                      __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                                      const int* __restrict__ lines, const int* __restrict__ positions,  // [B], [B*T]
                                      const __nv_bfloat16* __restrict__ q_norm, const __nv_bfloat16* __restrict__ k_norm,
                                      float eps, int T, int nq, int nkv, int L, int S) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HALF = D / 2;
  const int bt = blockIdx.x, head = blockIdx.y, i = threadIdx.x;  // i in [0, HALF)
  const int b = bt / T;
  const __nv_bfloat16* src = qkv + (size_t)bt * (nq + 2 * nkv) * D + (size_t)head * D;
  float x1 = __bfloat162float(src[i]), x2 = __bfloat162float(src[i + HALF]);
  const bool is_q = head < nq, is_k = !is_q && head < nq + nkv;
  if (is_q || is_k) {
    const __nv_bfloat16* nw = is_q ? q_norm : k_norm;
    if (nw != nullptr) {
      __shared__ float sred[HALF / 32];
      float ss = warp_sum(x1 * x1 + x2 * x2);
      if ((i & 31) == 0) sred[i >> 5] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < HALF / 32; ++w) tot += sred[w];
      const float rstd = rsqrtf(tot / (float)D + eps);
      // norm output is rounded to bf16 before the rotation, like the unfused module path
      x1 = __bfloat162float(__float2bfloat16(x1 * rstd * __bfloat162float(nw[i])));
      x2 = __bfloat162float(__float2bfloat16(x2 * rstd * __bfloat162float(nw[i + HALF])));
    }
    const float c = cos[(size_t)bt * HALF + i], s = sin[(size_t)bt * HALF + i];
    const float o1 = x1 * c - x2 * s, o2 = x2 * c + x1 * s;
    x1 = o1;
    x2 = o2;
  }
  if (is_q) {
    __nv_bfloat16* dst = q_out + ((size_t)bt * nq + head) * D;
    dst[i] = __float2bfloat16(x1);
    dst[i + HALF] = __float2bfloat16(x2);
    return;
  }
  const int kvh = is_k ? head - nq : head - nq - nkv;
  if (k_out != nullptr) {   // prefill: attention consumes the fresh (rotated) k / v directly, no cache read
    __nv_bfloat16* o = (is_k ? k_out : v_out) + ((size_t)bt * nkv + kvh) * D;
    o[i] = __float2bfloat16(x1);
    o[i + HALF] = __float2bfloat16(x2);
  }
  const int line = lines[b], pos = positions[bt];
  if (line < 0 || line >= L || pos < 0 || pos >= S) return;  // masked sequence / padding: skip the write
  __nv_bfloat16* cache = is_k ? k_cache : v_cache;
  __nv_bfloat16* dst = cache + (((size_t)line * nkv + kvh) * S + pos) * D;
  dst[i] = __float2bfloat16(x1);
  dst[i + HALF] = __float2bfloat16(x2);
}

void rope_kv_append_launch(const void* qkv, const float* cos, const float* sin, void* q_out, void* k_cache, void* v_cache,
                           const int* lines, const int* positions, const void* q_norm, const void* k_norm, float eps, int B,
                           int T, int nq, int nkv, int D, int L, int S, cudaStream_t stream, void* k_out, void* v_out) {
  dim3 grid(B * T, nq + 2 * nkv);
#define LAUNCH(DD)                                                                                                   \
  launch_pdl(rope_kv_append_kernel<DD>, grid, dim3(DD / 2), 0, stream, reinterpret_cast<const __nv_bfloat16*>(qkv), cos, \
             sin, reinterpret_cast<__nv_bfloat16*>(q_out), reinterpret_cast<__nv_bfloat16*>(k_out),                   \
             reinterpret_cast<__nv_bfloat16*>(v_out), reinterpret_cast<__nv_bfloat16*>(k_cache),                      \
             reinterpret_cast<__nv_bfloat16*>(v_cache), lines, positions, reinterpret_cast<const __nv_bfloat16*>(q_norm), \
             reinterpret_cast<const __nv_bfloat16*>(k_norm), eps, T, nq, nkv, L, S)
  if (D == 128) LAUNCH(128);
  else if (D == 64) LAUNCH(64);
  else if (D == 256) LAUNCH(256);
  else throw std::runtime_error("rope_kv_append: head_dim must be 64, 128 or 256");
#undef LAUNCH
}

// ---- plain appends --------------------------------------------------------------------------------------
// k_new/v_new [B,T,H,D] -> cache [L,H,S,D] at (lines[b], positions[b,t]); 16-byte vectors.
__global__ void kv_append_kernel(const uint4* __restrict__ k_new, const uint4* __restrict__ v_new, uint4* __restrict__ k_cache,
                                 uint4* __restrict__ v_cache, const int* __restrict__ lines, const int* __restrict__ positions,
                                 int T, int H, int DV, int L, int S, int total) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int dv = idx % DV, h = (idx / DV) % H, bt = idx / (DV * H);
  const int line = lines[bt / T], pos = positions[bt];
  if (line < 0 || line >= L || pos < 0 || pos >= S) return;
  const size_t dst = (((size_t)line * H + h) * S + pos) * DV + dv;
  k_cache[dst] = k_new[idx];
  v_cache[dst] = v_new[idx];
}

void kv_append_launch(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* lines,
                      const int* positions, int B, int T, int H, int row_bytes, int L, int S, cudaStream_t stream) {
  const int DV = row_bytes / 16, total = B * T * H * DV;
  launch_pdl(kv_append_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, reinterpret_cast<const uint4*>(k_new),
             reinterpret_cast<const uint4*>(v_new), reinterpret_cast<uint4*>(k_cache), reinterpret_cast<uint4*>(v_cache), lines,
             positions, T, H, DV, L, S, total);
}

// paged: cache [num_blocks, block_size, H, D]; slot_mapping [B*T] (-1 = skip)
__global__ void paged_kv_append_kernel(const uint4* __restrict__ k_new, const uint4* __restrict__ v_new, uint4* __restrict__ k_cache,
                                       uint4* __restrict__ v_cache, const int* __restrict__ slots, int HDV, int n_slots, int total) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int tok = idx / HDV, r = idx % HDV;
  const int slot = slots[tok];
  if (slot < 0 || slot >= n_slots) return;
  k_cache[(size_t)slot * HDV + r] = k_new[idx];
  v_cache[(size_t)slot * HDV + r] = v_new[idx];
}

void paged_kv_append_launch(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* slots, int ntok,
                            int tok_bytes, int n_slots, cudaStream_t stream) {
  const int HDV = tok_bytes / 16, total = ntok * HDV;
  launch_pdl(paged_kv_append_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, reinterpret_cast<const uint4*>(k_new),
             reinterpret_cast<const uint4*>(v_new), reinterpret_cast<uint4*>(k_cache), reinterpret_cast<uint4*>(v_cache), slots,
             HDV, n_slots, total);
}

}  // namespace nxdi
