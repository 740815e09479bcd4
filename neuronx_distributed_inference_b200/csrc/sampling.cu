// On-device sampling kernels: split arg-max over the vocabulary and global-top-k / top-p / temperature
// multinomial sampling (reference K8 cumsum + K9 topk/argmax kernels, modules/generation/sampling.py:241-464).
#include <cfloat>

#include "api.h"
#include "common.cuh"

namespace nxdi {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- arg-max -------------------------------------------------------------------------------------------
// grid (nsplit, B).  Partial (value,index) per CTA -> workspace; the last CTA of a row (atomic ticket) reduces.
// Vocabulary-sharded lm_head (sy.world > 1): the row's winner is exchanged with every peer INSIDE this kernel — the
// {value, tag} / {global index, tag} pairs go straight into each rank's slot over NVLink (LL stores), every rank polls its
// `world` slots and picks the same global winner (ties -> lowest index, like a flat arg-max).  Replaces local arg-max +
// NCCL all-gather + five torch glue kernels in the decode graph (reference: modules/generation/sampling.py:306-326).
template <typename T>
__global__ void __launch_bounds__(256) argmax_kernel(const T* __restrict__ logits, int64_t* __restrict__ out,
                                                     float* __restrict__ ws_val, int* __restrict__ ws_idx,
                                                     unsigned* __restrict__ tickets, int V, int ld, const ArgmaxSymm sy) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x, tid = threadIdx.x;
  const int chunk = (V + nsplit - 1) / nsplit;
  const int beg = split * chunk, end = min(V, beg + chunk);
  const T* row = logits + (size_t)b * ld;
  float best = -FLT_MAX;
  int bi = 0x7fffffff;
  for (int i = beg + tid; i < end; i += 256) {
    const float v = to_f(row[i]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = best; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    ws_val[b * nsplit + split] = best;
    ws_idx[b * nsplit + split] = bi;
    __threadfence();
    last = (atomicAdd(&tickets[b], 1u) == (unsigned)(nsplit - 1));
  }
  __syncthreads();
  if (last && tid < 32) {
    __threadfence();
    float v = -FLT_MAX;
    int ix = 0x7fffffff;
    for (int s = tid; s < nsplit; s += 32) {
      const float ov = reinterpret_cast<volatile float*>(ws_val)[b * nsplit + s];
      const int oi = reinterpret_cast<volatile int*>(ws_idx)[b * nsplit + s];
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    if (sy.world > 1) {
      const uint32_t tag = ll_tag(sy.step, sy.call);
      const int gix = ix + sy.rank * V;
      if (tid < sy.world) {
        float* dst = sy.slots[0];
#pragma unroll
        for (int d = 1; d < SYMM_MAX_RANKS; ++d)
          if (d == tid) dst = sy.slots[d];
        dst += (((size_t)sy.parity * sy.rows_max + b) * sy.world + sy.rank) * 4;
        st_ll(dst, v, tag);
        st_ll(dst + 2, __int_as_float(gix), tag);
      }
      v = -FLT_MAX;
      ix = 0x7fffffff;
      if (tid < sy.world) {
        const float* src = sy.slots[0];
#pragma unroll
        for (int d = 1; d < SYMM_MAX_RANKS; ++d)
          if (d == sy.rank) src = sy.slots[d];
        src += (((size_t)sy.parity * sy.rows_max + b) * sy.world + tid) * 4;
        float a, c;
        uint32_t fa, fc;
        const long long t0 = clock64();
        while (true) {
          ld_ll(src, a, fa);
          ld_ll(src + 2, c, fc);
          if (fa == tag && fc == tag) break;
          if (clock64() - t0 > 8000000000LL) {
            printf("argmax exchange: rank %d timed out waiting for rank %d (row %d, tag %u, seen %u %u)\n", sy.rank, tid, b, tag, fa, fc);
            __trap();
          }
        }
        v = a;
        ix = __float_as_int(c);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
      }
    }
    if (tid == 0) {
      out[b] = ix;
      tickets[b] = 0;  // ready for the next launch / graph replay
    }
  }
}

void argmax_launch(const void* logits, int dtype, int64_t* out, float* ws_val, int* ws_idx, unsigned* tickets, int B, int V,
                   int ld, int nsplit, const ArgmaxSymm* symm, cudaStream_t stream) {
  dim3 grid(nsplit, B);
  ArgmaxSymm sy{};
  if (symm != nullptr) sy = *symm;
  if (dtype == 0)
    launch_pdl(argmax_kernel<float>, grid, dim3(256), 0, stream, reinterpret_cast<const float*>(logits), out, ws_val, ws_idx,
               tickets, V, ld, sy);
  else
    launch_pdl(argmax_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(logits), out,
               ws_val, ws_idx, tickets, V, ld, sy);
}

// ---- top-k / top-p / temperature sampling ------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

constexpr int TOPK_MAX = 256;
constexpr int TOPK_THREADS = 1024;

// One CTA per row: 4-pass radix select of the K-th largest logit, gather the K winners, bitonic sort them
// (value desc, index asc), then one warp applies top-k mask / temperature / softmax / top-p / inverse-CDF.
template <typename T>
__global__ void __launch_bounds__(TOPK_THREADS) topk_sample_kernel(const T* __restrict__ logits, const int* __restrict__ top_k,
                                                                   const float* __restrict__ top_p, const float* __restrict__ temperature,
                                                                   const float* __restrict__ rand, int64_t* __restrict__ out, int V, int ld,
                                                                   int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ unsigned hist[256];
  __shared__ unsigned long long cand[TOPK_MAX];
  __shared__ unsigned s_prefix, s_need, s_count_gt, s_count_eq;
  __shared__ float s_p[TOPK_MAX];
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* row = logits + (size_t)b * ld;
  K = min(K, V);
  if (tid == 0) { s_prefix = 0; s_need = K; }
  __syncthreads();
  // radix select, most significant byte first
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < V; i += TOPK_THREADS) {
      const unsigned k = order_key(to_f(row[i]));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_need, acc = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (acc + hist[d] >= need) break;
        acc += hist[d];
      }
      s_prefix = prefix | ((unsigned)d << shift);
      s_need = need - acc;
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;  // key of the K-th largest element
  if (tid == 0) { s_count_gt = 0; s_count_eq = 0; }
  for (int i = tid; i < TOPK_MAX; i += TOPK_THREADS) cand[i] = 0ull;
  __syncthreads();
  const unsigned n_eq_take = s_need;  // how many elements equal to thr belong to the top-K
  for (int i = tid; i < V; i += TOPK_THREADS) {
    const unsigned k = order_key(to_f(row[i]));
    if (k > thr) {
      const unsigned slot = atomicAdd(&s_count_gt, 1u);
      if (slot < (unsigned)K) cand[slot] = ((unsigned long long)k << 32) | (unsigned)(0xffffffffu - (unsigned)i);
    }
  }
  __syncthreads();
  const unsigned n_gt = min(s_count_gt, (unsigned)K);
  for (int i = tid; i < V; i += TOPK_THREADS) {
    const unsigned k = order_key(to_f(row[i]));
    if (k == thr) {
      const unsigned slot = atomicAdd(&s_count_eq, 1u);
      if (slot < n_eq_take && n_gt + slot < (unsigned)K)
        cand[n_gt + slot] = ((unsigned long long)k << 32) | (unsigned)(0xffffffffu - (unsigned)i);
    }
  }
  __syncthreads();
  // bitonic sort (descending) of TOPK_MAX 64-bit keys; empty slots are 0 and sink to the end
  for (int size = 2; size <= TOPK_MAX; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < TOPK_MAX / 2) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = cand[lo], c = cand[hi];
        if ((a < c) == desc) { cand[lo] = c; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (tid < 32) {
    const int lane = tid;
    int k_eff = top_k[b];
    k_eff = (k_eff <= 0) ? K : min(k_eff, K);
    const float temp = temperature[b];
    const bool greedy = (temp == 0.f) || (k_eff == 1);
    const float inv_t = (temp == 0.f) ? 1.f : 1.f / temp;
    const float vmax = key_to_float((uint32_t)(cand[0] >> 32)) * inv_t;
    constexpr int PER = TOPK_MAX / 32;
    float e[PER];
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = lane * PER + j;
      float v = 0.f;
      if (i < k_eff) v = __expf(key_to_float((uint32_t)(cand[i] >> 32)) * inv_t - vmax);
      e[j] = v;
      lsum += v;
    }
    const float total = warp_sum(lsum);
    // inclusive scan of lane sums
    float incl = lsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    float run = (incl - lsum) / total;  // exclusive cumulative prob before my chunk
    const float tp = top_p[b];
    float kept = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const float pj = e[j] / total;
      const bool keep = run < tp;  // (cum - p) < top_p : always keeps the first token
      e[j] = keep ? pj : 0.f;
      kept += e[j];
      run += pj;
    }
    const float ktot = warp_sum(kept);
    float kincl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, kincl, o);
      if (lane >= o) kincl += n;
    }
    float cdf = (kincl - kept) / ktot;
    const float r = rand[b];
    int below = 0, nkept = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      cdf += e[j] / ktot;
      if (e[j] > 0.f) {
        ++nkept;
        if (cdf < r) ++below;
      }
    }
    int choice = (int)warp_sum((float)below);
    nkept = (int)warp_sum((float)nkept);
    choice = min(choice, max(nkept, 1) - 1);
    if (greedy) choice = 0;
    if (lane == 0) out[b] = (int64_t)(0xffffffffu - (unsigned)(cand[choice] & 0xffffffffull));
    (void)s_p;
  }
}

void topk_sample_launch(const void* logits, int dtype, const int* top_k, const float* top_p, const float* temperature,
                        const float* rand, int64_t* out, int B, int V, int ld, int K, cudaStream_t stream) {
  if (dtype == 0)
    launch_pdl(topk_sample_kernel<float>, dim3(B), dim3(TOPK_THREADS), 0, stream, reinterpret_cast<const float*>(logits), top_k,
               top_p, temperature, rand, out, V, ld, K);
  else
    launch_pdl(topk_sample_kernel<__nv_bfloat16>, dim3(B), dim3(TOPK_THREADS), 0, stream,
               reinterpret_cast<const __nv_bfloat16*>(logits), top_k, top_p, temperature, rand, out, V, ld, K);
}

}  // namespace nxdi
