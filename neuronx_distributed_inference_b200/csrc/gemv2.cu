// GEMV v2: TMA (3-D tensor map, UTMALDG) mbarrier-pipelined, stream-K balanced weight-streaming skinny GEMM
//          Y[T<=8, N] = f( rmsnorm(X)[T,K] · W[N,K]^T )      (+ fused one-shot all-reduce over NVLink)
//
// Why a second design (v1 = gemv.cuh, register-staged LDG): ncu on v1 showed warps >90 % stalled on
// long-scoreboard with DRAM at 51-64 % — in-flight bytes were bounded by registers (2 stages x 8 LDG.128 per
// warp) and a 16-row tile granularity left 15-35 % of the chip idle in the last wave.  v2 fixes both:
//   * ONE elected producer thread streams W with TMA tensor loads (cp.async.bulk.tensor.3d, completion on an
//     mbarrier): W[N,K] is described as a 3-D tensor {64 k, K/64 groups, N rows}; one instruction fetches the box
//     {64, 4, 16} = 16 weight rows x 256 k = 8 KB into a 128B-swizzled stage; up to 24 stages (192 KB) are in
//     flight per SM, independent of registers.  (A first version issued sixteen 512-byte 1-D bulk copies per
//     stage and was TMA-issue bound at 1.4 TB/s — see profiles/.)
//   * 8 consumer warps wait on the stage's full-barrier, take their 32-k slice as mma.sync A fragments straight
//     from shared memory (the 128B swizzle keeps the 8-lane LDS.128 phases bank-conflict free), multiply with
//     the token fragments of X (RMS-normalised in the prologue, bf16 in shared memory) and release the stage;
//   * stream-K: the flattened (tile, k-chunk) space is cut into gridDim equal contiguous ranges, so every SM
//     streams the same number of bytes; a tile that straddles CTAs is finished by the last arriver (atomic
//     ticket) which sums the partials in slot order — deterministic, no float atomics;
//   * the producer starts before griddepcontrol.wait (weights never depend on the previous kernel);
//   * fused all-reduce (MODE 1): LL protocol — every partial travels as an 8-byte {value, flag} store straight
//     into each peer's receive slot over NVLink; the receiver polls the same 8 bytes, so there is no separate
//     flag, no fence.sys round trip and no CTA barrier on the critical path.  Slots are self-resetting and
//     double-buffered by call parity (see parallel/symm.py).
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int G2_CONSUMER_WARPS = 8;
constexpr int G2_THREADS = (G2_CONSUMER_WARPS + 1) * 32;  // + producer warp
constexpr int G2_KC = 256;                                 // k elements per stage
constexpr int G2_STAGE_BYTES = 16 * G2_KC * 2;             // 8192, 128B-swizzled [16 rows][4 groups][64 k]
constexpr int G2_MAX_STAGES = 24;
constexpr int G2_SMEM_BUDGET = 224 * 1024;  // dynamic; leaves room for the few static __shared__ words

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// element offset (bf16) of 16-byte chunk `chunk` (0..7) of 128-byte line `line` inside a 128B-swizzled stage
__device__ __forceinline__ int swz128(int line, int chunk) { return line * 64 + ((chunk ^ (line & 7)) << 3); }
__device__ __forceinline__ void st_ll(float* p, float v, uint32_t flag) {  // 8-byte {value, flag}: single-copy atomic
  asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(flag) : "memory");
}
__device__ __forceinline__ void ld_ll(const float* p, float& v, uint32_t& flag) {
  uint32_t a, b;
  asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
  v = __uint_as_float(a);
  flag = b;
}

struct Gemv2Params {
  CUtensorMap tmap;   // W as {64, K/64, N}, box {64, 4, 16 (8 for GLU)}, SWIZZLE_128B
  GemvParams g;
  float* ws_part;     // [n_tiles * p_max][128] fp32 stream-K partials
  unsigned* tickets;  // [n_tiles]
  int p_max;
  int n_stages;
  int whole_tiles;  // 1: CTAs own whole 16-row tiles (no stream-K fix-up); chosen for small weights where the fix-up
                    // round trips (partials + ticket + re-read) cost more than the tile-count imbalance
};

template <bool GLU, int MODE>
__global__ void __launch_bounds__(G2_THREADS, 2) gemv2_kernel(const __grid_constant__ Gemv2Params pp) {
  const GemvParams& p = pp.g;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, N = p.N, T = p.T;
  const int NS = pp.n_stages;
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(p.w);
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* BIAS = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  const __nv_bfloat16* RES = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);

  // shared memory carve-up (stage_base 1024-aligned for the 128B swizzle):
  //   [stages][8 KB] | xs[T][2K+64] | red[8][128] f32 | rstd[64] f32 | barriers
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xs_stride = K * 2 + 64;
  uint8_t* xs = stage_base + (size_t)NS * G2_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(xs + (size_t)T * xs_stride);
  float* rstd_s = red + G2_CONSUMER_WARPS * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(rstd_s + 64);
  uint64_t* empty_bar = full_bar + G2_MAX_STAGES;
  __shared__ int s_flag;

  const int n_chunks = K / G2_KC;
  const int n_tiles = GLU ? ((N >> 1) + 7) >> 3 : (N + 15) >> 4;
  const long long U = (long long)n_tiles * n_chunks;
  const int G = gridDim.x, c = blockIdx.x;
  const long long u_beg = pp.whole_tiles ? ((long long)n_tiles * c / G) * n_chunks : (U * c) / G;
  const long long u_end = pp.whole_tiles ? ((long long)n_tiles * (c + 1) / G) * n_chunks : (U * (c + 1)) / G;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == G2_CONSUMER_WARPS) {
    // =========================== producer: one elected thread drives the TMA ===========================
    pdl_launch_dependents();
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&pp.tmap) : "memory");
      int tile = (int)(u_beg / n_chunks), chunk = (int)(u_beg % n_chunks), stage = 0;
      uint32_t ph = 0;
      const int count = (int)(u_end - u_beg);
      for (int i = 0; i < count; ++i) {
        mbar_wait(&empty_bar[stage], ph ^ 1u);
        mbar_expect_tx(&full_bar[stage], G2_STAGE_BYTES);
        uint8_t* dst = stage_base + (size_t)stage * G2_STAGE_BYTES;
        if (GLU) {  // 8 gate rows then 8 up rows (rows past the end are zero-filled by the TMA unit)
          tma_load_3d(dst, &pp.tmap, 0, chunk * 4, tile * 8, &full_bar[stage]);
          tma_load_3d(dst + G2_STAGE_BYTES / 2, &pp.tmap, 0, chunk * 4, (N >> 1) + tile * 8, &full_bar[stage]);
        } else {
          tma_load_3d(dst, &pp.tmap, 0, chunk * 4, tile * 16, &full_bar[stage]);
        }
        if (++chunk == n_chunks) { chunk = 0; ++tile; }
        if (++stage == NS) { stage = 0; ph ^= 1u; }
      }
    }
#ifdef NXDI_GEMV_PREFETCH_NEXT
    // ---- warm L2 for the NEXT skinny GEMM of the stream ----
    // Consecutive decode GEMVs cannot overlap (one 200 KB CTA per SM), so HBM idles for the ~5 us of kernel tail + launch +
    // x prologue between them.  The producer is done issuing its own loads one ring ahead of the consumers: from here it
    // asks L2 to fetch exactly what CTA c of the next kernel will request first (its first ring fill, 16-row x 256-col
    // units of 512-byte row segments); the requests outlive this kernel.  Spread over the 32 lanes.
    __syncwarp();  // lanes 1..31 wait here until lane 0 has issued the last of this kernel's own loads
    if (p.pf_w != nullptr) {
      const int K2 = p.pf_K, N2 = p.pf_N;
      const int n_chunks2 = K2 / G2_KC;
      const int n_tiles2 = p.pf_glu ? ((N2 >> 1) + 7) >> 3 : (N2 + 15) >> 4;
      const long long U2 = (long long)n_tiles2 * n_chunks2;
      const long long G2 = U2 / 4 < (long long)G ? (U2 / 4 > 0 ? U2 / 4 : 1) : (long long)G;   // gemv2_grid() of the next launch
      if ((long long)c < G2) {
        const long long b2 = (U2 * c) / G2, e2 = (U2 * (c + 1)) / G2;
        const int cnt = (int)((e2 - b2) < (long long)G2_MAX_STAGES ? (e2 - b2) : (long long)G2_MAX_STAGES);
        const char* W2 = reinterpret_cast<const char*>(p.pf_w);
        for (int i = 0; i < cnt; ++i) {
          const long long u2 = b2 + i;
          const int tile2 = (int)(u2 / n_chunks2), chunk2 = (int)(u2 % n_chunks2);
          // 16 row segments of 512 bytes (GLU: 8 gate rows + 8 up rows)
          if (lane < 16) {
            int row;
            if (p.pf_glu) row = lane < 8 ? tile2 * 8 + lane : (N2 >> 1) + tile2 * 8 + (lane - 8);
            else row = tile2 * 16 + lane;
            if (row < N2) {
              const char* a = W2 + ((size_t)row * K2 + (size_t)chunk2 * G2_KC) * 2;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(512) : "memory");
            }
          }
        }
      }
    }
#endif
    return;
  }

  // ================================= consumer warps =================================
  const int g = lane >> 2, t4 = lane & 3;
  const int ctid = tid;  // 0..255
  pdl_wait();
  // ---- X prologue: x * gamma -> bf16 in shared memory, ONE pass; the per-token 1/rms is a scalar, so it is applied to the
  //      fp32 accumulators in the epilogue instead of to x (saves the second pass over x and one CTA barrier).
  //      All T rows are fetched in one flattened loop: the loads are independent, one L2 latency in total.
  {
    float ss[GEMV_MAX_T];
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) ss[t] = 0.f;
    const int nvec = K >> 3;
    const bool has_norm = p.norm_w != nullptr;
    const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
    const float o = p.norm_offset;
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      if (t >= T) break;
      const uint4* src = reinterpret_cast<const uint4*>(X + (size_t)t * p.ldx);
      uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
      float acc = 0.f;
      for (int v = ctid; v < nvec; v += 256) {
        uint4 q = ldg_cached(src + v);
        if (has_norm) {
          acc += bf16lo(q.x) * bf16lo(q.x) + bf16hi(q.x) * bf16hi(q.x) + bf16lo(q.y) * bf16lo(q.y) +
                 bf16hi(q.y) * bf16hi(q.y) + bf16lo(q.z) * bf16lo(q.z) + bf16hi(q.z) * bf16hi(q.z) +
                 bf16lo(q.w) * bf16lo(q.w) + bf16hi(q.w) * bf16hi(q.w);
          const uint4 gm = ldg_cached(gw + v);
          q.x = pack_bf16(bf16lo(q.x) * (bf16lo(gm.x) + o), bf16hi(q.x) * (bf16hi(gm.x) + o));
          q.y = pack_bf16(bf16lo(q.y) * (bf16lo(gm.y) + o), bf16hi(q.y) * (bf16hi(gm.y) + o));
          q.z = pack_bf16(bf16lo(q.z) * (bf16lo(gm.z) + o), bf16hi(q.z) * (bf16hi(gm.z) + o));
          q.w = pack_bf16(bf16lo(q.w) * (bf16lo(gm.w) + o), bf16hi(q.w) * (bf16hi(gm.w) + o));
        }
        dst[v] = q;
      }
      ss[t] = acc;
    }
    if (has_norm) {
#pragma unroll
      for (int t = 0; t < GEMV_MAX_T; ++t) {
        if (t < T) {
          float v = warp_sum(ss[t]);
          if (lane == 0) rstd_s[warp * 8 + t] = v;
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  // 1/rms of token `col` (valid after the barrier above; read in the epilogue)
  auto rstd_of = [&](int col) -> float {
    if (p.norm_w == nullptr) return 1.f;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < G2_CONSUMER_WARPS; ++w) tot += rstd_s[w * 8 + col];
    return rsqrtf(tot / (float)K + p.eps);
  };

  const bool tok_ok = g < T;
  const uint8_t* xrow = xs + (size_t)g * xs_stride;
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};

  // epilogue of one finished 16x8 tile whose fp32 sums are in `vals` (thread ctid<128 owns (col=ctid>>4,row=ctid&15))
  auto finalize = [&](int tile, float v_gate_or_val, float v_up) {
    const int col = ctid >> 4, row = ctid & 15;
    if (col >= T) return;
    if (GLU) {
      if (row >= 8) return;
      const int n = tile * 8 + row, half = N >> 1;
      if (n >= half) return;
      const float rs = rstd_of(col);
      float gate = v_gate_or_val * rs, up = v_up * rs;
      if (BIAS != nullptr) {
        gate += __bfloat162float(BIAS[n]);
        up += __bfloat162float(BIAS[half + n]);
      }
      const float a = p.act == 1 ? silu(gate) : (p.act == 2 ? gelu_tanh(gate) : gelu_erf(gate));
      Y[(size_t)col * p.ldy + n] = __float2bfloat16(a * up);
    } else {
      const int n = tile * 16 + row;
      if (n >= N) return;
      float v = v_gate_or_val * rstd_of(col);
      if (MODE == 0) {
        if (BIAS != nullptr) v += __bfloat162float(BIAS[n]);
        if (RES != nullptr) v += __bfloat162float(RES[(size_t)col * p.ldy + n]);
        Y[(size_t)col * p.ldy + n] = __float2bfloat16(v);
      } else {
        // LL all-reduce phase 1: {value, 1} straight into every peer's slot (including mine)
        const SymmArgs& s = p.symm;
        const size_t off = (((size_t)(s.parity * s.world + s.rank) * 8 + col) * s.n_max + n) * 2;
#pragma unroll
        for (int d = 0; d < SYMM_MAX_RANKS; ++d)
          if (d < s.world) st_ll(s.recv[d] + off, v, 1u);
      }
    }
  };

  // Each consumer warp owns whole stages (local unit i -> warp i % 8): 8 k-iterations of 32 per stage, one
  // barrier wait and one release per 8 KB.  A tile's units are spread over the warps; the flush joins them.
  const int count = (int)(u_end - u_beg);
  int cur_tile = (int)(u_beg / n_chunks);
  int chunk_first = (int)(u_beg % n_chunks);  // chunk index of local unit `seg_beg`
  int seg_beg = 0;                             // local index of the first unit of the current tile segment
  long long tile_u0 = u_beg;
  while (seg_beg < count) {
    const int seg_len = min(n_chunks - chunk_first, count - seg_beg);
    const long long u = u_beg + seg_beg + seg_len - 1;  // last unit of this segment (global)
    for (int i = seg_beg + ((warp - seg_beg) & 7); i < seg_beg + seg_len; i += G2_CONSUMER_WARPS) {
      const int chunk = chunk_first + (i - seg_beg);
      const int stage = i % NS;
      const uint32_t ph = (uint32_t)((i / NS) & 1);
      mbar_wait(&full_bar[stage], ph);
      const __nv_bfloat16* sA = reinterpret_cast<const __nv_bfloat16*>(stage_base + (size_t)stage * G2_STAGE_BYTES);
      const uint8_t* xk = xrow + (size_t)(chunk * G2_KC + t4 * 8) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kg = j >> 1, ch = ((j & 1) << 2) + t4;
        const uint4 a0 = *reinterpret_cast<const uint4*>(sA + swz128(g * 4 + kg, ch));
        const uint4 a1 = *reinterpret_cast<const uint4*>(sA + swz128((g + 8) * 4 + kg, ch));
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (tok_ok) xv = *reinterpret_cast<const uint4*>(xk + j * 64);
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
      }
      // the warp-collective MMAs have consumed every lane's fragments: the stage can be refilled
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }
    {
      const int chunk = chunk_first + seg_len - 1;
    // ---- flush: cross-warp reduce of the 16x8 tile ----
    float* r = red + warp * 128;
    r[g * 8 + 2 * t4] = c0[0] + c1[0];
    r[g * 8 + 2 * t4 + 1] = c0[1] + c1[1];
    r[(g + 8) * 8 + 2 * t4] = c0[2] + c1[2];
    r[(g + 8) * 8 + 2 * t4 + 1] = c0[3] + c1[3];
#pragma unroll
    for (int q = 0; q < 4; ++q) c0[q] = c1[q] = 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const long long t_first = (long long)cur_tile * n_chunks, t_last = t_first + n_chunks;  // unit range of the tile
    const bool whole = (tile_u0 == t_first) && (u + 1 == t_last);
    float va = 0.f, vb = 0.f;
    if (ctid < 128) {
      const int col = ctid >> 4, row = ctid & 15;
#pragma unroll
      for (int w = 0; w < G2_CONSUMER_WARPS; ++w) {
        va += red[w * 128 + row * 8 + col];
        if (GLU) vb += red[w * 128 + ((row + 8) & 15) * 8 + col];
      }
    }
    if (whole) {
      if (ctid < 128) finalize(cur_tile, va, vb);
    } else {
      // stream-K: this CTA owns only part of the tile.  slot = my index among the CTAs that cover it.
      long long cf = (t_first * G) / U;
      while (((U * (cf + 1)) / G) <= t_first) ++cf;
      while (((U * cf) / G) > t_first) --cf;
      long long cl = ((t_last - 1) * G) / U;
      while (((U * (cl + 1)) / G) <= t_last - 1) ++cl;
      while (((U * cl) / G) > t_last - 1) --cl;
      const int slot = (int)(c - cf), n_parts = (int)(cl - cf + 1);
      float* my = pp.ws_part + ((size_t)cur_tile * pp.p_max + slot) * 128;
      if (ctid < 128) {
        const int col = ctid >> 4, row = ctid & 15;
        my[row * 8 + col] = va;  // natural (row, col) layout: rows 0-7 gate / 8-15 up for GLU tiles
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ctid == 0) s_flag = (atomicAdd(&pp.tickets[cur_tile], 1u) == (unsigned)(n_parts - 1)) ? 1 : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_flag) {
        __threadfence();
        if (ctid < 128) {
          const int col = ctid >> 4, row = ctid & 15;
          float sa = 0.f, sb = 0.f;
          for (int q = 0; q < n_parts; ++q) {
            const float* pq = pp.ws_part + ((size_t)cur_tile * pp.p_max + q) * 128;
            sa += __ldcg(pq + row * 8 + col);
            if (GLU) sb += __ldcg(pq + ((row + 8) & 15) * 8 + col);
          }
          finalize(cur_tile, sa, sb);
        }
        if (ctid == 0) pp.tickets[cur_tile] = 0;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // red / s_flag reusable
    }
    cur_tile += 1;
    tile_u0 = u + 1;
    seg_beg += seg_len;
    chunk_first = 0;
  }

  if (MODE == 1) {
    // ---- LL all-reduce phase 2: every CTA polls a strided share of ALL (column, token) slots ----
    const SymmArgs& s = p.symm;
    float* my_recv = s.recv[0];
#pragma unroll
    for (int d = 1; d < SYMM_MAX_RANKS; ++d)
      if (d == s.rank) my_recv = s.recv[d];
    const int total = T * N;
    for (int e = c * 256 + ctid; e < total; e += G * 256) {
      const int col = e / N, n = e % N;
      float v = 0.f;
      for (int r = 0; r < s.world; ++r) {
        float* slot = my_recv + (((size_t)(s.parity * s.world + r) * 8 + col) * s.n_max + n) * 2;
        float x;
        uint32_t f;
        const long long t0 = clock64();
        while (true) {
          ld_ll(slot, x, f);
          if (f != 0u) break;
          if (clock64() - t0 > 8000000000LL) {
            printf("gemv_allreduce: rank %d timed out waiting for rank %d (col %d n %d)\n", s.rank, r, col, n);
            __trap();
          }
        }
        st_ll(slot, 0.f, 0u);  // self-reset: this slot is reused two collectives from now
        v += x;
      }
      if (BIAS != nullptr) v += __bfloat162float(BIAS[n]);
      if (RES != nullptr) v += __bfloat162float(RES[(size_t)col * p.ldy + n]);
      Y[(size_t)col * p.ldy + n] = __float2bfloat16(v);
    }
  }
}

static int g2_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static size_t g2_fixed_smem(int T, int K) {
  return (size_t)T * (K * 2 + 64) + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + 2 * G2_MAX_STAGES * sizeof(uint64_t) +
         128 + 1024;  // + worst-case 1 KB alignment slack
}

bool gemv2_supported(int T, int K) {
  return K % G2_KC == 0 && g2_fixed_smem(T, K) + 8 * G2_STAGE_BYTES <= G2_SMEM_BUDGET;
}

// Small weights: whole tiles per CTA (no cross-CTA fix-up).  Large weights: stream-K (perfect byte balance matters more).
static bool g2_whole_tiles(int N, int K, bool glu) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NXDI_B200_GEMV_WHOLE_TILES");
    mode = e ? atoi(e) : 2;   // 0 never, 1 always, 2 heuristic
  }
  if (mode != 2) return mode == 1;
  const int n_tiles = glu ? ((N / 2) + 7) / 8 : (N + 15) / 16;
  const int sms = g2_num_sms();
  const int active = std::min(sms, n_tiles);
  const int per = (n_tiles + active - 1) / active;            // tiles of the busiest CTA
  const double tile_bytes = 16.0 * K * 2;
  // one SM pulls at most ~100 GB/s; all of them together ~6 TB/s (bytes per microsecond below)
  const double bw_sm = std::min(100e3, 6.0e6 / active);
  const double t_whole = per * tile_bytes / bw_sm;
  // stream-K: perfectly balanced bytes + the fix-up (partials, ticket, re-read: ~3.5-4 us; tools/bench_gemv_fixed.py:
  // 4096x4096 11.1 -> 6.8 us, 1536x4096 10.4 -> 4.6 us once it is gone; Llama-8B decode 3.67 -> 3.19 ms/step)
  const double t_streamk = (double)N * K * 2 / 6.0e6 + 4.0;
  return t_whole <= t_streamk + 0.5;
}

int gemv2_grid(int N, int K, bool glu) {
  const int n_tiles = glu ? ((N / 2) + 7) / 8 : (N + 15) / 16;
  if (g2_whole_tiles(N, K, glu)) return std::min(g2_num_sms(), n_tiles);
  const long long U = (long long)n_tiles * (K / G2_KC);
  return (int)std::min<long long>(g2_num_sms(), std::max<long long>(U / 4, 1));
}

int gemv2_pmax(int N, int K, bool glu) {
  const int n_tiles = glu ? ((N / 2) + 7) / 8 : (N + 15) / 16;
  return gemv2_grid(N, K, glu) / n_tiles + 3;
}

template <bool GLU, int MODE>
static void launch_gemv2(const Gemv2Params& pp, cudaStream_t stream) {
  auto kern = gemv2_kernel<GLU, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BUDGET);
    configured = true;
  }
  const size_t smem = g2_fixed_smem(pp.g.T, pp.g.K) + (size_t)pp.n_stages * G2_STAGE_BYTES;
  const int grid = gemv2_grid(pp.g.N, pp.g.K, GLU);
  launch_pdl(kern, dim3(grid), dim3(G2_THREADS), smem, stream, pp);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// W [N, K] bf16 row-major as a 3-D tensor {64, K/64, N}; box {64, 4, rows}
void make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K, int box_rows) {
  cuuint64_t gdim[3] = {64, (cuuint64_t)(K / 64), (cuuint64_t)N};
  cuuint64_t gstride[2] = {128, (cuuint64_t)K * 2};
  cuuint32_t box[3] = {64, 4, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
}

void gemv2_launch(const GemvParams& p, int mode, float* ws_part, unsigned* tickets, cudaStream_t stream) {
  Gemv2Params pp;
  pp.g = p;
  make_weight_tmap(&pp.tmap, p.w, p.N, p.K, p.act != 0 ? 8 : 16);
  pp.ws_part = ws_part;
  pp.tickets = tickets;
  const bool glu = p.act != 0;
  pp.p_max = gemv2_pmax(p.N, p.K, glu);
  pp.whole_tiles = g2_whole_tiles(p.N, p.K, glu) ? 1 : 0;
  const size_t fixed = g2_fixed_smem(p.T, p.K);
  // Ring depth: a MULTIPLE OF 8 stages.  Consumer warp w owns the units i == w (mod 8); with NS % 8 == 0 it always
  // meets the same stages, lap after lap, so its mbarrier phase parity is unambiguous (with NS = 10 a warp could
  // wait for lap 1 of a stage whose lap 0 had not completed yet and fall through: found on the TP=2 decode shapes).
  static int budget_kb = -1;
  if (budget_kb < 0) {
    const char* e = getenv("NXDI_B200_GEMV_SMEM_KB");
    budget_kb = e ? atoi(e) : 224;
  }
  const size_t budget = std::min<size_t>((size_t)budget_kb * 1024, G2_SMEM_BUDGET);
  int ns = budget > fixed ? (int)((budget - fixed) / G2_STAGE_BYTES) : 0;
  ns = std::min(ns, G2_MAX_STAGES) / 8 * 8;
  pp.n_stages = std::max(ns, 8);
  if (mode == 1) launch_gemv2<false, 1>(pp, stream);
  else if (glu) launch_gemv2<true, 0>(pp, stream);
  else launch_gemv2<false, 0>(pp, stream);
}


// =====================================================================================================================
// GEMV CHAIN: up to 4 dependent skinny GEMMs in ONE persistent launch (one CTA per SM, grid barriers in between).
//
//   decode layer tail:   A  h1  = attn_out · Wo^T  (+ all-reduce) + h
//                        B  u   = swiglu( rmsnorm(h1) · Wgu^T )
//                        C  h2  = u · Wd^T        (+ all-reduce) + h1
//                        D  qkv = rmsnorm(h2) · Wqkv(next layer)^T + b          (or the lm_head for the last layer)
//
// Why: at decode the four GEMVs of a layer are separate launches whose fixed cost (launch gap, x prologue, pipeline fill,
// stream-K tail: ~7-10 us each) dwarfs their streaming time once the weights are sharded (TP4: 12 us of bytes, 70 us of
// wall clock per layer).  Inside one kernel the TMA producer never stops: weights do not depend on activations, so while
// the consumer warps sit in the grid barrier and rebuild x for the next phase the producer is already filling the ring
// (192 KB/SM = ~28 MB chip-wide, ~4 us of HBM time) with the NEXT phase's weights.  A grid barrier costs ~2 us.
// The body of a phase is gemv2_kernel's (same stage format, stream-K split, ticketed fix-up, LL all-reduce); GLU / MODE
// are runtime flags here, the stage ring and its mbarrier phases run on across phase boundaries.
constexpr int CHAIN_MAX_PHASES = 4;

struct ChainPhase {
  CUtensorMap tmap;
  GemvParams g;
  float* ws_part;
  unsigned* tickets;
  int p_max, glu, mode, pad_;
};
struct ChainParams {
  ChainPhase ph[CHAIN_MAX_PHASES];
  unsigned* bar;  // [0] arrival count, [32] generation (separate 128-byte lines: pollers must not queue with arrivals)
  int n_phases, n_stages, xs_bytes, pad_;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(G2_THREADS, 1) gemv_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NS = cp.n_stages;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = stage_base + (size_t)NS * G2_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(xs + cp.xs_bytes);
  float* rstd_s = red + G2_CONSUMER_WARPS * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(rstd_s + 64);
  uint64_t* empty_bar = full_bar + G2_MAX_STAGES;
  __shared__ int s_flag;
  const int G = gridDim.x, c = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == G2_CONSUMER_WARPS) {
    // ======================= producer: streams the weights of ALL phases back to back =======================
    pdl_launch_dependents();
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (int pi = 0; pi < cp.n_phases; ++pi) {
        const ChainPhase& P = cp.ph[pi];
        asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap) : "memory");
        const int K = P.g.K, N = P.g.N;
        const int n_chunks = K / G2_KC;
        const int n_tiles = P.glu ? ((N >> 1) + 7) >> 3 : (N + 15) >> 4;
        const long long U = (long long)n_tiles * n_chunks;
        const long long u_beg = (U * c) / G, u_end = (U * (c + 1)) / G;
        int tile = (int)(u_beg / n_chunks), chunk = (int)(u_beg % n_chunks);
        const int count = (int)(u_end - u_beg);
        for (int i = 0; i < count; ++i) {
          mbar_wait(&empty_bar[stage], ph ^ 1u);
          mbar_expect_tx(&full_bar[stage], G2_STAGE_BYTES);
          uint8_t* dst = stage_base + (size_t)stage * G2_STAGE_BYTES;
          if (P.glu) {
            tma_load_3d(dst, &P.tmap, 0, chunk * 4, tile * 8, &full_bar[stage]);
            tma_load_3d(dst + G2_STAGE_BYTES / 2, &P.tmap, 0, chunk * 4, (N >> 1) + tile * 8, &full_bar[stage]);
          } else {
            tma_load_3d(dst, &P.tmap, 0, chunk * 4, tile * 16, &full_bar[stage]);
          }
          if (++chunk == n_chunks) { chunk = 0; ++tile; }
          if (++stage == NS) { stage = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ================================= consumer warps =================================
  const int g = lane >> 2, t4 = lane & 3;
  const int ctid = tid;  // 0..255
  pdl_wait();
  unsigned my_gen = 0;
  if (ctid == 0) my_gen = ld_acquire_gpu(cp.bar + 32);
  long long ibase = 0;  // running unit index of this CTA across phases (stage ring position)

  for (int pi = 0; pi < cp.n_phases; ++pi) {
    const ChainPhase& P = cp.ph[pi];
    const GemvParams& p = P.g;
    const bool GLU = P.glu != 0;
    const int MODE = P.mode;
    const int K = p.K, N = p.N, T = p.T;
    const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const __nv_bfloat16* BIAS = reinterpret_cast<const __nv_bfloat16*>(p.bias);
    const __nv_bfloat16* RES = reinterpret_cast<const __nv_bfloat16*>(p.residual);
    __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);
    const int xs_stride = K * 2 + 64;
    const int n_chunks = K / G2_KC;
    const int n_tiles = GLU ? ((N >> 1) + 7) >> 3 : (N + 15) >> 4;
    const long long U = (long long)n_tiles * n_chunks;
    const long long u_beg = (U * c) / G, u_end = (U * (c + 1)) / G;

    // ---- X prologue (inputs of phases > 0 were written by other SMs in this launch: L2 loads only) ----
    {
      const int nvec = K >> 3;
      for (int t = 0; t < T; ++t) {
        const uint4* src = reinterpret_cast<const uint4*>(X + (size_t)t * p.ldx);
        uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
        float acc = 0.f;
        for (int v = ctid; v < nvec; v += 256) {
          uint4 q = __ldcg(src + v);
          dst[v] = q;
          acc += bf16lo(q.x) * bf16lo(q.x) + bf16hi(q.x) * bf16hi(q.x) + bf16lo(q.y) * bf16lo(q.y) +
                 bf16hi(q.y) * bf16hi(q.y) + bf16lo(q.z) * bf16lo(q.z) + bf16hi(q.z) * bf16hi(q.z) +
                 bf16lo(q.w) * bf16lo(q.w) + bf16hi(q.w) * bf16hi(q.w);
        }
        if (p.norm_w != nullptr) {
          acc = warp_sum(acc);
          if (lane == 0) rstd_s[warp * 8 + t] = acc;
        }
      }
      if (p.norm_w != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int t = 0; t < T; ++t) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < G2_CONSUMER_WARPS; ++w) tot += rstd_s[w * 8 + t];
          const float rstd = rsqrtf(tot / (float)K + p.eps);
          uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
          const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
          const float o = p.norm_offset;
          for (int v = ctid; v < nvec; v += 256) {
            uint4 q = dst[v];
            uint4 gm = ldg_cached(gw + v);
            q.x = pack_bf16(bf16lo(q.x) * rstd * (bf16lo(gm.x) + o), bf16hi(q.x) * rstd * (bf16hi(gm.x) + o));
            q.y = pack_bf16(bf16lo(q.y) * rstd * (bf16lo(gm.y) + o), bf16hi(q.y) * rstd * (bf16hi(gm.y) + o));
            q.z = pack_bf16(bf16lo(q.z) * rstd * (bf16lo(gm.z) + o), bf16hi(q.z) * rstd * (bf16hi(gm.z) + o));
            q.w = pack_bf16(bf16lo(q.w) * rstd * (bf16lo(gm.w) + o), bf16hi(q.w) * rstd * (bf16hi(gm.w) + o));
            dst[v] = q;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }

    const bool tok_ok = g < T;
    const uint8_t* xrow = xs + (size_t)g * xs_stride;
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};

    auto finalize = [&](int tile, float v_gate_or_val, float v_up) {
      const int col = ctid >> 4, row = ctid & 15;
      if (col >= T) return;
      if (GLU) {
        if (row >= 8) return;
        const int n = tile * 8 + row, half = N >> 1;
        if (n >= half) return;
        float gate = v_gate_or_val, up = v_up;
        if (BIAS != nullptr) {
          gate += __bfloat162float(BIAS[n]);
          up += __bfloat162float(BIAS[half + n]);
        }
        const float a = p.act == 1 ? silu(gate) : (p.act == 2 ? gelu_tanh(gate) : gelu_erf(gate));
        Y[(size_t)col * p.ldy + n] = __float2bfloat16(a * up);
      } else {
        const int n = tile * 16 + row;
        if (n >= N) return;
        float v = v_gate_or_val;
        if (MODE == 0) {
          if (BIAS != nullptr) v += __bfloat162float(BIAS[n]);
          if (RES != nullptr) v += bf16lo((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(RES) + (size_t)col * p.ldy + n));
          Y[(size_t)col * p.ldy + n] = __float2bfloat16(v);
        } else {
          const SymmArgs& s = p.symm;
          const size_t off = (((size_t)(s.parity * s.world + s.rank) * 8 + col) * s.n_max + n) * 2;
#pragma unroll
          for (int d = 0; d < SYMM_MAX_RANKS; ++d)
            if (d < s.world) st_ll(s.recv[d] + off, v, 1u);
        }
      }
    };

    const int count = (int)(u_end - u_beg);
    int cur_tile = (int)(u_beg / n_chunks);
    int chunk_first = (int)(u_beg % n_chunks);
    int seg_beg = 0;
    long long tile_u0 = u_beg;
    while (seg_beg < count) {
      const int seg_len = min(n_chunks - chunk_first, count - seg_beg);
      const long long u = u_beg + seg_beg + seg_len - 1;
      // local unit i has ring index ibase + i; warp w owns ring indices == w (mod 8)
      for (int i = seg_beg + (int)((warp - (ibase + seg_beg)) & 7); i < seg_beg + seg_len; i += G2_CONSUMER_WARPS) {
        const int chunk = chunk_first + (i - seg_beg);
        const long long ri = ibase + i;
        const int stage = (int)(ri % NS);
        const uint32_t ph = (uint32_t)((ri / NS) & 1);
        mbar_wait(&full_bar[stage], ph);
        const __nv_bfloat16* sA = reinterpret_cast<const __nv_bfloat16*>(stage_base + (size_t)stage * G2_STAGE_BYTES);
        const uint8_t* xk = xrow + (size_t)(chunk * G2_KC + t4 * 8) * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kg = j >> 1, ch = ((j & 1) << 2) + t4;
          const uint4 a0 = *reinterpret_cast<const uint4*>(sA + swz128(g * 4 + kg, ch));
          const uint4 a1 = *reinterpret_cast<const uint4*>(sA + swz128((g + 8) * 4 + kg, ch));
          uint4 xv = make_uint4(0u, 0u, 0u, 0u);
          if (tok_ok) xv = *reinterpret_cast<const uint4*>(xk + j * 64);
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
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      {
        float* r = red + warp * 128;
        r[g * 8 + 2 * t4] = c0[0] + c1[0];
        r[g * 8 + 2 * t4 + 1] = c0[1] + c1[1];
        r[(g + 8) * 8 + 2 * t4] = c0[2] + c1[2];
        r[(g + 8) * 8 + 2 * t4 + 1] = c0[3] + c1[3];
#pragma unroll
        for (int q = 0; q < 4; ++q) c0[q] = c1[q] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const long long t_first = (long long)cur_tile * n_chunks, t_last = t_first + n_chunks;
        const bool whole = (tile_u0 == t_first) && (u + 1 == t_last);
        float va = 0.f, vb = 0.f;
        if (ctid < 128) {
          const int col = ctid >> 4, row = ctid & 15;
#pragma unroll
          for (int w = 0; w < G2_CONSUMER_WARPS; ++w) {
            va += red[w * 128 + row * 8 + col];
            if (GLU) vb += red[w * 128 + ((row + 8) & 15) * 8 + col];
          }
        }
        if (whole) {
          if (ctid < 128) finalize(cur_tile, va, vb);
        } else {
          long long cf = (t_first * G) / U;
          while (((U * (cf + 1)) / G) <= t_first) ++cf;
          while (((U * cf) / G) > t_first) --cf;
          long long cl = ((t_last - 1) * G) / U;
          while (((U * (cl + 1)) / G) <= t_last - 1) ++cl;
          while (((U * cl) / G) > t_last - 1) --cl;
          const int slot = (int)(c - cf), n_parts = (int)(cl - cf + 1);
          float* my = P.ws_part + ((size_t)cur_tile * P.p_max + slot) * 128;
          if (ctid < 128) {
            const int col = ctid >> 4, row = ctid & 15;
            my[row * 8 + col] = va;
          }
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (ctid == 0) s_flag = (atomicAdd(&P.tickets[cur_tile], 1u) == (unsigned)(n_parts - 1)) ? 1 : 0;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (s_flag) {
            __threadfence();
            if (ctid < 128) {
              const int col = ctid >> 4, row = ctid & 15;
              float sa = 0.f, sb = 0.f;
              for (int q = 0; q < n_parts; ++q) {
                const float* pq = P.ws_part + ((size_t)cur_tile * P.p_max + q) * 128;
                sa += __ldcg(pq + row * 8 + col);
                if (GLU) sb += __ldcg(pq + ((row + 8) & 15) * 8 + col);
              }
              finalize(cur_tile, sa, sb);
            }
            if (ctid == 0) P.tickets[cur_tile] = 0;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      cur_tile += 1;
      tile_u0 = u + 1;
      seg_beg += seg_len;
      chunk_first = 0;
    }
    ibase += count;

    if (MODE == 1) {
      const SymmArgs& s = p.symm;
      float* my_recv = s.recv[0];
#pragma unroll
      for (int d = 1; d < SYMM_MAX_RANKS; ++d)
        if (d == s.rank) my_recv = s.recv[d];
      const int total = T * N;
      for (int e = c * 256 + ctid; e < total; e += G * 256) {
        const int col = e / N, n = e % N;
        float v = 0.f;
        for (int r = 0; r < s.world; ++r) {
          float* slot = my_recv + (((size_t)(s.parity * s.world + r) * 8 + col) * s.n_max + n) * 2;
          float x;
          uint32_t f;
          const long long t0 = clock64();
          while (true) {
            ld_ll(slot, x, f);
            if (f != 0u) break;
            if (clock64() - t0 > 8000000000LL) {
              printf("gemv_chain: rank %d timed out waiting for rank %d (phase %d col %d n %d)\n", s.rank, r, pi, col, n);
              __trap();
            }
          }
          st_ll(slot, 0.f, 0u);
          v += x;
        }
        if (BIAS != nullptr) v += __bfloat162float(BIAS[n]);
        if (RES != nullptr) v += bf16lo((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(RES) + (size_t)col * p.ldy + n));
        Y[(size_t)col * p.ldy + n] = __float2bfloat16(v);
      }
    }

    // ---- grid barrier: the next phase reads what every SM just wrote ----
    if (pi + 1 < cp.n_phases) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ctid == 0) {
        __threadfence();
        const unsigned old = atomicAdd(cp.bar, 1u);
        if (old == (unsigned)(G - 1)) {
          atomicExch(cp.bar, 0u);
          __threadfence();
          atomicAdd(cp.bar + 32, 1u);
        } else {
          const long long t0 = clock64();
          while (ld_acquire_gpu(cp.bar + 32) == my_gen) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000LL) {  // ~2 s: a grid that is not co-resident would spin forever
              printf("gemv_chain: grid barrier timeout (cta %d phase %d)\n", c, pi);
              __trap();
            }
          }
        }
        ++my_gen;
        __threadfence();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
}

// xs region + fixed part for a chain whose widest input has K_max columns
static size_t chain_fixed_smem(int T, int K_max) {
  return (size_t)T * (K_max * 2 + 64) + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + 2 * G2_MAX_STAGES * sizeof(uint64_t) + 128 +
         1024;
}

bool gemv_chain_supported(int T, int K_max) { return chain_fixed_smem(T, K_max) + 8 * G2_STAGE_BYTES <= G2_SMEM_BUDGET; }

size_t gemv_chain_ws_floats(const GemvParams* ph, int n) {
  size_t tot = 0;
  const int G = g2_num_sms();
  for (int i = 0; i < n; ++i) {
    const bool glu = ph[i].act != 0;
    const int n_tiles = glu ? ((ph[i].N / 2) + 7) / 8 : (ph[i].N + 15) / 16;
    tot += (size_t)n_tiles * (G / n_tiles + 3) * 128;
  }
  return tot;
}

size_t gemv_chain_tickets(const GemvParams* ph, int n) {
  size_t tot = 0;
  for (int i = 0; i < n; ++i) tot += ph[i].act != 0 ? ((ph[i].N / 2) + 7) / 8 : (ph[i].N + 15) / 16;
  return tot;
}

void gemv_chain_launch(const GemvParams* ph, const int* modes, int n, float* ws_part, unsigned* tickets, unsigned* bar,
                       cudaStream_t stream) {
  if (n < 1 || n > CHAIN_MAX_PHASES) throw std::runtime_error("gemv_chain: 1..4 phases");
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  const int G = g2_num_sms();
  int K_max = 0;
  for (int i = 0; i < n; ++i) {
    ChainPhase& P = cp.ph[i];
    P.g = ph[i];
    const bool glu = ph[i].act != 0;
    make_weight_tmap(&P.tmap, ph[i].w, ph[i].N, ph[i].K, glu ? 8 : 16);
    const int n_tiles = glu ? ((ph[i].N / 2) + 7) / 8 : (ph[i].N + 15) / 16;
    P.p_max = G / n_tiles + 3;
    P.glu = glu ? 1 : 0;
    P.mode = modes[i];
    P.ws_part = ws_part;
    P.tickets = tickets;
    ws_part += (size_t)n_tiles * P.p_max * 128;
    tickets += n_tiles;
    K_max = std::max(K_max, ph[i].K);
    if (ph[i].K % G2_KC != 0) throw std::runtime_error("gemv_chain: K must be a multiple of 256");
  }
  cp.bar = bar;
  cp.n_phases = n;
  const size_t fixed = chain_fixed_smem(ph[0].T, K_max);
  int ns = G2_SMEM_BUDGET > fixed ? (int)((G2_SMEM_BUDGET - fixed) / G2_STAGE_BYTES) : 0;
  ns = std::min(ns, G2_MAX_STAGES) / 8 * 8;
  if (ns < 8) throw std::runtime_error("gemv_chain: activations do not fit in shared memory");
  cp.n_stages = ns;
  cp.xs_bytes = ph[0].T * (K_max * 2 + 64);
  auto kern = gemv_chain_kernel;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BUDGET);
    configured = true;
  }
  const size_t smem = fixed + (size_t)ns * G2_STAGE_BYTES;
  launch_pdl(kern, dim3(G), dim3(G2_THREADS), smem, stream, cp);
}

}  // namespace nxdi
