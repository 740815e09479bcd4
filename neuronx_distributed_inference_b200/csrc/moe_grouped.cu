// Mixture-of-experts prefill: on-device token permutation around the grouped tcgen05 GEMM (gemm_tcgen05.cu, grp_tile_expert).
//
//   plan     one CTA: histogram of the (token, slot) entries routed to the experts this rank owns, exclusive scan with every
//            expert's segment padded to whole 128-row tiles, tile -> expert table, entry -> permuted row (`pos`) and its inverse
//   gather   permuted activations xp[row] = x[token(row)] (optionally x affinity: Llama-4 scales the expert INPUT), zero pad rows
//   GEMM 1   h  = act(xp @ w_gate_up[e]^T)        grouped, GLU epilogue
//   GEMM 2   y  = h @ w_down[e]^T                 grouped
//   combine  out[token] = sum_slot affinity * y[pos[token, slot]]      (fp32 accumulate, fixed slot order: deterministic)
//
// Everything is sized by the static bound R = roundup128(N * k + E_local * 127): no host synchronisation, no data-dependent
// launch shapes — the whole layer replays inside a CUDA graph whatever the router decides.  The reference gets the same effect
// from its blockwise matmul over a padded token->block table (modules/moe_v2.py:25-133 and the blockwise kernels behind
// ExpertMLPsV2); tokens of one expert are contiguous here so that one TMA box feeds one UMMA tile.
#include <stdexcept>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int MOE_MAX_E = 512;
constexpr int MOE_PLAN_THREADS = 1024;

struct MoePlanArgs {
  const int* topk_i;   // [entries] global expert ids
  int entries;         // N * k
  int expert_offset, e_local;
  int R;               // static row bound (multiple of 128)
  int* pos;            // [entries] permuted row of the entry, -1 when its expert lives on another rank
  int* row_entry;      // [R] entry held by a permuted row, -1 for padding
  int* tile_expert;    // [R / 128] local expert of a row tile, -1 when unused
};

__global__ void __launch_bounds__(MOE_PLAN_THREADS) moe_plan_kernel(const MoePlanArgs a) {
  __shared__ int count[MOE_MAX_E], start[MOE_MAX_E + 1], cursor[MOE_MAX_E];
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  for (int e = tid; e < a.e_local; e += MOE_PLAN_THREADS) count[e] = cursor[e] = 0;
  pdl_wait();   // row_entry / tile_expert / pos may still be read by the previous layer's kernels
  for (int r = tid; r < a.R; r += MOE_PLAN_THREADS) a.row_entry[r] = -1;
  for (int t = tid; t < a.R / 128; t += MOE_PLAN_THREADS) a.tile_expert[t] = -1;
  __syncthreads();
  for (int i = tid; i < a.entries; i += MOE_PLAN_THREADS) {
    const int e = a.topk_i[i] - a.expert_offset;
    if (e >= 0 && e < a.e_local) atomicAdd(&count[e], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int s = 0;
    for (int e = 0; e < a.e_local; ++e) {
      start[e] = s;
      s += (count[e] + 127) & ~127;
    }
    start[a.e_local] = s;
  }
  __syncthreads();
  for (int e = tid; e < a.e_local; e += MOE_PLAN_THREADS)
    for (int t = start[e] >> 7; t < (start[e + 1] >> 7); ++t) a.tile_expert[t] = e;
  for (int i = tid; i < a.entries; i += MOE_PLAN_THREADS) {
    const int e = a.topk_i[i] - a.expert_offset;
    int r = -1;
    if (e >= 0 && e < a.e_local) {
      r = start[e] + atomicAdd(&cursor[e], 1);
      a.row_entry[r] = i;
    }
    a.pos[i] = r;
  }
}

// one CTA per permuted row
__global__ void __launch_bounds__(128) moe_gather_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ topk_w,
                                                         const int* __restrict__ row_entry, const int* __restrict__ tile_expert,
                                                         __nv_bfloat16* __restrict__ xp, int H, int k, int scale_input) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  if (tile_expert[r >> 7] < 0) return;
  const int ent = row_entry[r];
  uint4* dst = reinterpret_cast<uint4*>(xp + (size_t)r * H);
  if (ent < 0) {
    for (int v = threadIdx.x; v < H / 8; v += 128) dst[v] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)(ent / k) * H);
  if (!scale_input) {
    for (int v = threadIdx.x; v < H / 8; v += 128) dst[v] = __ldg(src + v);
  } else {
    const float w = topk_w[ent];
    for (int v = threadIdx.x; v < H / 8; v += 128) {
      const uint4 q = __ldg(src + v);
      dst[v] = make_uint4(pack_bf16(bf16lo(q.x) * w, bf16hi(q.x) * w), pack_bf16(bf16lo(q.y) * w, bf16hi(q.y) * w),
                          pack_bf16(bf16lo(q.z) * w, bf16hi(q.z) * w), pack_bf16(bf16lo(q.w) * w, bf16hi(q.w) * w));
    }
  }
}

// one CTA per token
__global__ void __launch_bounds__(128) moe_combine_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ topk_w,
                                                          const int* __restrict__ pos, __nv_bfloat16* __restrict__ out, int H, int k,
                                                          int scale_input) {
  pdl_launch_dependents();
  pdl_wait();
  const int tok = blockIdx.x;
  __shared__ int s_pos[64];
  __shared__ float s_w[64];
  if ((int)threadIdx.x < k) {
    s_pos[threadIdx.x] = pos[(size_t)tok * k + threadIdx.x];
    s_w[threadIdx.x] = scale_input ? 1.f : topk_w[(size_t)tok * k + threadIdx.x];
  }
  __syncthreads();
  for (int v = threadIdx.x; v < H / 8; v += 128) {
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < k; ++s) {
      const int r = s_pos[s];
      if (r < 0) continue;
      const float w = s_w[s];
      const uint4 q = ldg_act(y + (size_t)r * H + v * 8);
      f[0] += w * bf16lo(q.x); f[1] += w * bf16hi(q.x); f[2] += w * bf16lo(q.y); f[3] += w * bf16hi(q.y);
      f[4] += w * bf16lo(q.z); f[5] += w * bf16hi(q.z); f[6] += w * bf16lo(q.w); f[7] += w * bf16hi(q.w);
    }
    *reinterpret_cast<uint4*>(out + (size_t)tok * H + v * 8) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

void moe_plan_launch(const int* topk_i, int entries, int expert_offset, int e_local, int R, int* pos, int* row_entry, int* tile_expert,
                     cudaStream_t stream) {
  if (e_local > MOE_MAX_E) throw std::runtime_error("moe_plan: more than 512 local experts");
  if (R % 128 != 0) throw std::runtime_error("moe_plan: row bound must be a multiple of 128");
  MoePlanArgs a{topk_i, entries, expert_offset, e_local, R, pos, row_entry, tile_expert};
  launch_pdl(moe_plan_kernel, dim3(1), dim3(MOE_PLAN_THREADS), 0, stream, a);
}

void moe_gather_launch(const void* x, const float* topk_w, const int* row_entry, const int* tile_expert, void* xp, int R, int H, int k,
                       int scale_input, cudaStream_t stream) {
  if (H % 8 != 0) throw std::runtime_error("moe_gather: hidden size must be a multiple of 8");
  launch_pdl(moe_gather_kernel, dim3(R), dim3(128), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), topk_w, row_entry, tile_expert,
             reinterpret_cast<__nv_bfloat16*>(xp), H, k, scale_input);
}

void moe_combine_launch(const void* y, const float* topk_w, const int* pos, void* out, int N, int H, int k, int scale_input,
                        cudaStream_t stream) {
  if (k > 64) throw std::runtime_error("moe_combine: top-k > 64");
  launch_pdl(moe_combine_kernel, dim3(N), dim3(128), 0, stream, reinterpret_cast<const __nv_bfloat16*>(y), topk_w, pos,
             reinterpret_cast<__nv_bfloat16*>(out), H, k, scale_input);
}

}  // namespace nxdi
