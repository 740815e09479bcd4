// GEMV v2: TMA (3-D tensor map, UTMALDG) mbarrier-pipelined weight-streaming skinny GEMM for decode
//          Y[T<=8, N] = f( rmsnorm(X)[T,K] · W[N,K]^T )      (+ fused one-shot all-reduce over NVLink)
//
// Design (round 2):
//   * ONE elected producer thread streams W with TMA tensor loads (cp.async.bulk.tensor.3d, completion on an
//     mbarrier): W[N,K] is described as a 3-D tensor {64 k, K/64 groups, N rows}; one instruction fetches the box
//     {64, 2, 16} = 16 weight rows x 128 k = 4 KB into a 128B-swizzled stage.
//   * CO-RESIDENCY: the CTA is sized to <= ~108 KB of shared memory so that TWO consecutive decode kernels fit on an SM.
//     Under programmatic dependent launch the NEXT kernel's producer then fills its ring (weights never depend on the
//     previous kernel) while THIS kernel is still computing: HBM no longer idles for the ~5 us launch / wait / x-prologue /
//     flush window between dependent GEMVs, which was what separated 3.17 ms/step from the 2.28 ms streaming floor at TP1 and
//     dominated the step at TP4/TP8 (profiles/decode_overheads_r1.md).
//   * 8 consumer warps; a stage always belongs to warp (stage & 7), so a warp meets its stages lap after lap in order and the
//     mbarrier phase parity is unambiguous for ANY ring depth >= 8.  A-fragments of mma.sync m16n8k16 come straight from the
//     swizzled stage (8-lane LDS.128 phases are conflict free); the token fragments of X (gamma folded in, bf16) sit in shared
//     memory; 1/rms is applied to the fp32 accumulators in the epilogue (single pass over x).
//   * stream-K for large weights (equal bytes per SM, ticketed deterministic fix-up), whole 16-row tiles per CTA for small ones.
//   * fused all-reduce (MODE 1): LL protocol — every partial travels as an 8-byte {value, tag} store straight into each peer's
//     receive slot over NVLink; the receiver polls all sources of an element IN PARALLEL (one round trip, not `world`).  The
//     tag is an epoch = (device-side step counter, call index in the step): a skipped / repeated collective can no longer be
//     mistaken for the expected one (it times out and traps instead), and slots need no reset.  Slots are double-buffered by
//     call parity (a rank can be at most one collective ahead of its slowest peer).
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int G2_CONSUMER_WARPS = 8;
constexpr int G2_THREADS = (G2_CONSUMER_WARPS + 1) * 32;  // + producer warp
constexpr int G2_KG = 8;                                   // 128-byte k groups per row per stage (one per consumer warp)
constexpr int G2_KC = 64 * G2_KG;                          // 512 k elements per stage
constexpr int G2_STAGE_BYTES = 16 * G2_KC * 2;             // 16384: [8 k groups][16 rows][64 k], 128B-swizzled
constexpr int G2_MAX_STAGES = 13;
constexpr int G2_SMEM_BUDGET = 224 * 1024;   // hard cap (dynamic); leaves room for the few static __shared__ words
constexpr int G2_SMEM_CORES = 108 * 1024;    // default budget: two such CTAs (or one + a 112 KB attention CTA) per SM

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
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Gemv2Params {
  CUtensorMap tmap;   // W as {64 k, N rows, K/64 groups}, box {64, 16 (8 for GLU / rows8), 8}, SWIZZLE_128B
  GemvParams g;
  float* ws_part;     // [n_tiles * p_max][128] fp32 stream-K partials
  unsigned* tickets;  // [n_tiles]
  int p_max;
  int n_stages;
  int rows8;        // 1 (plain epilogues, narrow N): 8-row tiles — twice the CTAs for a projection with < ~100 16-row tiles
                    // (rows 8..15 of the MMA are zeros; the tensor pipe is idle at decode anyway)
  unsigned long long* prof;  // debug timeline: [8] u64 per CTA (tools/prof_decode.py) or null
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
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* BIAS = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  const __nv_bfloat16* RES = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);

  // shared memory carve-up (stage_base 1024-aligned for the 128B swizzle):
  //   [stages][16 KB] | xs[T][2*Kp+64] | red[8][128] f32 | rstd[64] f32 | barriers        Kp = K rounded up to 512
  const int n_chunks = (K + G2_KC - 1) / G2_KC;
  const int Kp = n_chunks * G2_KC;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xs_stride = Kp * 2 + 64;
  uint8_t* xs = stage_base + (size_t)NS * G2_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(xs + (size_t)T * xs_stride);
  float* rstd_s = red + G2_CONSUMER_WARPS * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(rstd_s + 64);
  uint64_t* empty_bar = full_bar + G2_MAX_STAGES;
  __shared__ int s_flag;

  const bool rows8 = !GLU && pp.rows8 != 0;
  const int TR = rows8 ? 8 : 16;                       // output rows per tile (plain epilogues)
  const int n_tiles = GLU ? ((N >> 1) + 7) >> 3 : (N + TR - 1) / TR;
  const long long U = (long long)n_tiles * n_chunks;
  const int G = gridDim.x, c = blockIdx.x;
  const long long u_beg = pp.whole_tiles ? ((long long)n_tiles * c / G) * n_chunks : (U * c) / G;
  const long long u_end = pp.whole_tiles ? ((long long)n_tiles * (c + 1) / G) * n_chunks : (U * (c + 1)) / G;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], G2_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == G2_CONSUMER_WARPS) {
    // =========================== producer: one elected thread drives the TMA ===========================
    // The dependent kernel may start right away: it only prefetches ITS weights until this grid has completed.
    pdl_launch_dependents();
    // One issuing thread sustains one copy per ~150-230 ns whatever its size (tools/bench_stream.cu: 4 KB boxes 2.7 TB/s,
    // 8 KB 5.4 TB/s, 16 KB 7.2 TB/s chip-wide): the stage is therefore ONE 16 KB box (two 8 KB boxes for GLU tiles).
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&pp.tmap) : "memory");
      int tile = (int)(u_beg / n_chunks), chunk = (int)(u_beg % n_chunks), stage = 0;
      uint32_t ph = 0;
      const int count = (int)(u_end - u_beg);
      for (int i = 0; i < count; ++i) {
        mbar_wait(&empty_bar[stage], ph ^ 1u);
        mbar_expect_tx(&full_bar[stage], rows8 ? G2_STAGE_BYTES / 2 : G2_STAGE_BYTES);
        uint8_t* dst = stage_base + (size_t)stage * G2_STAGE_BYTES;
        // k groups past K/64 and rows past N are zero-filled by the TMA unit (and not fetched)
        if (GLU) {  // 8 gate rows, then 8 up rows
          tma_load_3d(dst, &pp.tmap, 0, tile * 8, chunk * G2_KG, &full_bar[stage]);
          tma_load_3d(dst + G2_STAGE_BYTES / 2, &pp.tmap, 0, (N >> 1) + tile * 8, chunk * G2_KG, &full_bar[stage]);
        } else {
          tma_load_3d(dst, &pp.tmap, 0, tile * TR, chunk * G2_KG, &full_bar[stage]);
        }
        if (++chunk == n_chunks) { chunk = 0; ++tile; }
        if (++stage == NS) { stage = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ================================= consumer warps =================================
  const int g = lane >> 2, t4 = lane & 3;
  const int ctid = tid;  // 0..255
  unsigned long long* prof = pp.prof ? pp.prof + (size_t)blockIdx.x * 8 : nullptr;
  if (prof && ctid == 0) { prof[0] = gtimer(); prof[1] = clock64(); }
  const int nvec = K >> 3;
  const bool has_norm = p.norm_w != nullptr;
  // ---- before the dependency resolves: everything that does not depend on the previous kernel ----
  // gamma -> parked in the (not yet used) row 0 of xs; zero the k padding of every x row (K..Kp)
  if (has_norm) {
    const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
    uint4* g0 = reinterpret_cast<uint4*>(xs);
    for (int v = ctid; v < nvec; v += 256) g0[v] = ldg_cached(gw + v);
  }
  if (Kp != K) {
    const int padv = (Kp - K) >> 3;
    for (int i = ctid; i < T * padv; i += 256)
      reinterpret_cast<uint4*>(xs + (size_t)(i / padv) * xs_stride)[nvec + i % padv] = make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();
  if (prof && ctid == 0) prof[2] = clock64();
  // all-reduce tag of this call: (device step counter << 8 | call index) + 1; the counter is bumped by the host-enqueued
  // begin_step op before the first collective of every forward (parallel/symm.py), so graph replays get fresh tags
  uint32_t ar_tag = 1u;
  if (MODE == 1) ar_tag = ll_tag(p.symm.step, p.symm.call);
  // ---- X prologue: x * gamma -> bf16 in shared memory, ONE pass; the per-token 1/rms is a scalar, so it is applied to the
  //      fp32 accumulators in the epilogue instead of to x.  The (vector, token) work items of a thread are flattened and
  //      fetched in batches of 8 independent 16-byte loads: ONE L2 round trip for T=2,K=4096.  Within a vector the tokens run
  //      from T-1 down to 0, so the gamma parked in row 0 is overwritten last (by the thread that owns that vector).
  {
    float ss[GEMV_MAX_T];
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) ss[t] = 0.f;
    const float o = p.norm_offset;
    const int per_thread = (nvec - ctid + 255) / 256;   // vectors of this thread
    const int items = per_thread > 0 ? per_thread * T : 0;
    for (int e0 = 0; e0 < items; e0 += 8) {
      uint4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j;
        if (e < items) {
          const int v = ctid + 256 * (e / T), t = T - 1 - e % T;
          q[j] = ldg_act(reinterpret_cast<const uint4*>(X + (size_t)t * p.ldx) + v);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j;
        if (e < items) {
          const int v = ctid + 256 * (e / T), t = T - 1 - e % T;
          uint4 w = q[j];
          if (has_norm) {
            const float sq = bf16lo(w.x) * bf16lo(w.x) + bf16hi(w.x) * bf16hi(w.x) + bf16lo(w.y) * bf16lo(w.y) +
                             bf16hi(w.y) * bf16hi(w.y) + bf16lo(w.z) * bf16lo(w.z) + bf16hi(w.z) * bf16hi(w.z) +
                             bf16lo(w.w) * bf16lo(w.w) + bf16hi(w.w) * bf16hi(w.w);
#pragma unroll
            for (int tt = 0; tt < GEMV_MAX_T; ++tt)
              if (tt == t) ss[tt] += sq;
            const uint4 gm = reinterpret_cast<const uint4*>(xs)[v];
            w.x = pack_bf16(bf16lo(w.x) * (bf16lo(gm.x) + o), bf16hi(w.x) * (bf16hi(gm.x) + o));
            w.y = pack_bf16(bf16lo(w.y) * (bf16lo(gm.y) + o), bf16hi(w.y) * (bf16hi(gm.y) + o));
            w.z = pack_bf16(bf16lo(w.z) * (bf16lo(gm.z) + o), bf16hi(w.z) * (bf16hi(gm.z) + o));
            w.w = pack_bf16(bf16lo(w.w) * (bf16lo(gm.w) + o), bf16hi(w.w) * (bf16hi(gm.w) + o));
          }
          reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride)[v] = w;
        }
      }
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
  if (prof && ctid == 0) prof[3] = clock64();
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

  // epilogue operands of the tile being streamed (bias / residual of this thread's (col, row)): fetched when the tile STARTS,
  // so their L2 latency hides behind the weight stream instead of sitting on the critical path after the last stage
  float pre_b0 = 0.f, pre_b1 = 0.f, pre_r = 0.f;
  auto prefetch_epilogue = [&](int tile) {
    pre_b0 = pre_b1 = pre_r = 0.f;
    if (ctid >= 128) return;
    const int col = ctid >> 4, row = ctid & 15;
    if (col >= T) return;
    if (GLU) {
      const int n = tile * 8 + row, half = N >> 1;
      if (row < 8 && n < half && BIAS != nullptr) {
        pre_b0 = __bfloat162float(BIAS[n]);
        pre_b1 = __bfloat162float(BIAS[half + n]);
      }
    } else if (MODE == 0) {
      const int n = tile * TR + row;
      if (row < TR && n < N) {
        if (BIAS != nullptr) pre_b0 = __bfloat162float(BIAS[n]);
        if (RES != nullptr) pre_r = ldg_act_bf16(RES + (size_t)col * p.ldy + n);
      }
    }
  };

  // epilogue of one finished 16x8 tile whose fp32 sums are in `vals` (thread ctid<128 owns (col=ctid>>4,row=ctid&15))
  auto finalize = [&](int tile, float v_gate_or_val, float v_up) {
    const int col = ctid >> 4, row = ctid & 15;
    if (col >= T) return;
    if (GLU) {
      if (row >= 8) return;
      const int n = tile * 8 + row, half = N >> 1;
      if (n >= half) return;
      const float rs = rstd_of(col);
      const float gate = v_gate_or_val * rs + pre_b0, up = v_up * rs + pre_b1;
      const float a = p.act == 1 ? silu(gate) : (p.act == 2 ? gelu_tanh(gate) : gelu_erf(gate));
      Y[(size_t)col * p.ldy + n] = __float2bfloat16(a * up);
    } else {
      if (row >= TR) return;
      const int n = tile * TR + row;
      if (n >= N) return;
      float v = v_gate_or_val * rstd_of(col);
      if (MODE == 0) {
        Y[(size_t)col * p.ldy + n] = __float2bfloat16(v + pre_b0 + pre_r);
      } else {
        // LL all-reduce phase 1: {value, tag} straight into every peer's slot (including mine)
        const SymmArgs& s = p.symm;
        const size_t off = (((size_t)(s.parity * s.world + s.rank) * 8 + col) * s.n_max + n) * 2;
#pragma unroll
        for (int d = 0; d < SYMM_MAX_RANKS; ++d)
          if (d < s.world) st_ll(s.recv[d] + off, v, ar_tag);
      }
    }
  };

  // MODE 1: the elements this thread will reduce after the exchange are known now — fetch their bias / residual early
  const int ar_total = (MODE == 1) ? T * N : 0;
  constexpr int AR_PRE = 4;
  float ar_pre[AR_PRE];
  if (MODE == 1) {
#pragma unroll
    for (int q = 0; q < AR_PRE; ++q) {
      const int e = c * 256 + ctid + q * G * 256;
      ar_pre[q] = 0.f;
      if (e < ar_total) {
        const int col = e / N, n = e % N;
        if (BIAS != nullptr) ar_pre[q] += __bfloat162float(BIAS[n]);
        if (RES != nullptr) ar_pre[q] += ldg_act_bf16(RES + (size_t)col * p.ldy + n);
      }
    }
  }

  // Every consumer warp takes part in every stage: warp w owns k group w (64 k = two MMA pairs) of the 512-k stage, so a
  // stage is drained by 8 warps at once and each warp meets the stages strictly in order (no mbarrier parity aliasing for any
  // ring depth).  The flush joins the warps' partial sums.
  const int count = (int)(u_end - u_beg);
  int cur_tile = (int)(u_beg / n_chunks);
  int chunk_first = (int)(u_beg % n_chunks);  // chunk index of local unit `seg_beg`
  int seg_beg = 0;                             // local index of the first unit of the current tile segment
  long long tile_u0 = u_beg;
  int stage = 0;
  uint32_t lap = 0;
  while (seg_beg < count) {
    const int seg_len = min(n_chunks - chunk_first, count - seg_beg);
    const long long u = u_beg + seg_beg + seg_len - 1;  // last unit of this segment (global)
    prefetch_epilogue(cur_tile);
    for (int i = 0; i < seg_len; ++i) {
      const int chunk = chunk_first + i;
      mbar_wait(&full_bar[stage], lap);
      const __nv_bfloat16* sA = reinterpret_cast<const __nv_bfloat16*>(stage_base + (size_t)stage * G2_STAGE_BYTES);
      const uint8_t* xk = xrow + (size_t)(chunk * G2_KC + warp * 64 + t4 * 8) * 2;
      // stage layout: line = kg * rows + row (rows = 16; GLU / rows8: two halves of [8 kg][8 rows])
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ch = (j << 2) + t4;
        uint4 a0, a1 = make_uint4(0u, 0u, 0u, 0u);
        if (GLU) {
          a0 = *reinterpret_cast<const uint4*>(sA + swz128(warp * 8 + g, ch));
          a1 = *reinterpret_cast<const uint4*>(sA + (G2_STAGE_BYTES / 4) + swz128(warp * 8 + g, ch));
        } else if (rows8) {
          a0 = *reinterpret_cast<const uint4*>(sA + swz128(warp * 8 + g, ch));
        } else {
          a0 = *reinterpret_cast<const uint4*>(sA + swz128(warp * 16 + g, ch));
          a1 = *reinterpret_cast<const uint4*>(sA + swz128(warp * 16 + g + 8, ch));
        }
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
      // the warp-collective MMAs have consumed every lane's fragments of this warp's k group
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == NS) { stage = 0; lap ^= 1u; }
    }
    {
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
  if (prof && ctid == 0) prof[4] = clock64();

  if (MODE == 1) {
    // ---- LL all-reduce phase 2: every CTA polls a strided share of ALL (column, token) slots; the `world` sources of an
    //      element are polled together (independent loads in flight: one L2 round trip per poll, not `world`) ----
    const SymmArgs& s = p.symm;
    float* my_recv = s.recv[0];
#pragma unroll
    for (int d = 1; d < SYMM_MAX_RANKS; ++d)
      if (d == s.rank) my_recv = s.recv[d];
    int q = 0;
    for (int e = c * 256 + ctid; e < ar_total; e += G * 256, ++q) {
      const int col = e / N, n = e % N;
      const float* slot0 = my_recv + (((size_t)(s.parity * s.world) * 8 + col) * s.n_max + n) * 2;
      const size_t rstride = (size_t)8 * s.n_max * 2;
      float x[SYMM_MAX_RANKS];
      const long long t0 = clock64();
      while (true) {
        uint32_t f[SYMM_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < SYMM_MAX_RANKS; ++r) {
          f[r] = ar_tag;
          x[r] = 0.f;
          if (r < s.world) ld_ll(slot0 + r * rstride, x[r], f[r]);
        }
        bool ok = true;
#pragma unroll
        for (int r = 0; r < SYMM_MAX_RANKS; ++r) ok = ok && (f[r] == ar_tag);
        if (ok) break;
        if (clock64() - t0 > 8000000000LL) {
          printf("gemv_allreduce: rank %d timed out (tag %u col %d n %d; seen %u %u %u %u %u %u %u %u)\n", s.rank, ar_tag, col, n,
                 f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
          __trap();
        }
      }
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < SYMM_MAX_RANKS; ++r) v += x[r];   // rank order: bitwise identical on every rank
      float extra = 0.f;
      if (q < AR_PRE) {
#pragma unroll
        for (int qq = 0; qq < AR_PRE; ++qq)
          if (qq == q) extra = ar_pre[qq];
      } else {
        if (BIAS != nullptr) extra += __bfloat162float(BIAS[n]);
        if (RES != nullptr) extra += ldg_act_bf16(RES + (size_t)col * p.ldy + n);
      }
      Y[(size_t)col * p.ldy + n] = __float2bfloat16(v + extra);
    }
  }
  if (prof && ctid == 0) {
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    prof[5] = clock64();
    prof[6] = gtimer();
    prof[7] = ((unsigned long long)count << 32) | sm;
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
  const int Kp = (K + G2_KC - 1) / G2_KC * G2_KC;
  return (size_t)T * (Kp * 2 + 64) + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + 2 * G2_MAX_STAGES * sizeof(uint64_t) +
         128 + 1024;  // + worst-case 1 KB alignment slack
}

bool gemv2_supported(int T, int K) {
  return K % 64 == 0 && g2_fixed_smem(T, K) + 3 * G2_STAGE_BYTES <= G2_SMEM_BUDGET;
}

// Small weights: whole tiles per CTA (no cross-CTA fix-up).  Large weights: stream-K (perfect byte balance matters more).
static bool g2_rows8(int N, bool glu) { return !glu && N % 8 == 0 && (N + 15) / 16 < 100; }
static int g2_ntiles(int N, bool glu) { return glu ? ((N / 2) + 7) / 8 : (g2_rows8(N, glu) ? (N + 7) / 8 : (N + 15) / 16); }
static double g2_tile_bytes(int N, int K, bool glu) { return (g2_rows8(N, glu) ? 8.0 : 16.0) * K * 2; }

static bool g2_whole_tiles(int N, int K, bool glu) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NXDI_B200_GEMV_WHOLE_TILES");
    mode = e ? atoi(e) : 2;   // 0 never, 1 always, 2 heuristic
  }
  if (mode != 2) return mode == 1;
  const int n_tiles = g2_ntiles(N, glu);
  const int sms = g2_num_sms();
  const int active = std::min(sms, n_tiles);
  const int per = (n_tiles + active - 1) / active;            // tiles of the busiest CTA
  const double tile_bytes = g2_tile_bytes(N, K, glu);
  // one SM pulls at most ~100 GB/s; all of them together ~6 TB/s (bytes per microsecond below)
  const double bw_sm = std::min(100e3, 6.0e6 / active);
  const double t_whole = per * tile_bytes / bw_sm;
  // stream-K: perfectly balanced bytes + the fix-up (partials, ticket, re-read: ~3.5-4 us; tools/bench_gemv_fixed.py:
  // 4096x4096 11.1 -> 6.8 us, 1536x4096 10.4 -> 4.6 us once it is gone; Llama-8B decode 3.67 -> 3.19 ms/step)
  const double t_streamk = (double)N * K * 2 / 6.0e6 + 4.0;
  return t_whole <= t_streamk + 0.5;
}

int gemv2_grid(int N, int K, bool glu) {
  const int n_tiles = g2_ntiles(N, glu);
  if (g2_whole_tiles(N, K, glu)) return std::min(g2_num_sms(), n_tiles);
  const long long U = (long long)n_tiles * ((K + G2_KC - 1) / G2_KC);
  return (int)std::min<long long>(g2_num_sms(), std::max<long long>(U / 2, 1));   // >= 32 KB of weights per CTA
}

int gemv2_ntiles(int N, bool glu) { return g2_ntiles(N, glu); }

int gemv2_pmax(int N, int K, bool glu) {
  const int n_tiles = g2_ntiles(N, glu);
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

// W [N, K] bf16 row-major as a 3-D tensor {64 k, N rows, K/64 groups} (strides 2 B, K*2 B, 128 B); box {64, rows, 8}:
// one instruction fetches rows x 512 k, and the box lands in shared memory as [k group][row][64 k] — 128-byte lines whose
// index & 7 is the row & 7, so the eight rows an LDS.128 phase touches sit in eight different swizzle positions.
void make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K, int box_rows) {
  cuuint64_t gdim[3] = {64, (cuuint64_t)N, (cuuint64_t)(K / 64)};
  cuuint64_t gstride[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)G2_KG};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
}

// Ring depth.  Default budget G2_SMEM_CORES keeps the CTA small enough for TWO decode kernels per SM (see the header);
// never deeper than the busiest CTA's unit count (a 2-tile CTA does not need 22 stages); at least 8 stages (one per warp).
// If even 8 stages do not fit the co-residency budget (wide K x many tokens) the kernel takes what it needs up to 224 KB.
static int g2_pick_stages(const GemvParams& p, bool glu, bool whole) {
  static int budget_kb = -1;
  if (budget_kb < 0) {
    const char* e = getenv("NXDI_B200_GEMV_SMEM_KB");
    budget_kb = e ? atoi(e) : G2_SMEM_CORES / 1024;
  }
  const size_t fixed = g2_fixed_smem(p.T, p.K);
  const size_t budget = std::min<size_t>((size_t)budget_kb * 1024, G2_SMEM_BUDGET);
  int ns = budget > fixed ? (int)((budget - fixed) / G2_STAGE_BYTES) : 0;
  const int n_tiles = g2_ntiles(p.N, glu);
  const int n_chunks = (p.K + G2_KC - 1) / G2_KC;
  const int G = gemv2_grid(p.N, p.K, glu);
  const long long per_cta = whole ? (long long)((n_tiles + G - 1) / G) * n_chunks : ((long long)n_tiles * n_chunks + G - 1) / G;
  ns = (int)std::min<long long>(ns, per_cta);
  ns = std::min(ns, G2_MAX_STAGES);
  if (ns < 3) {   // wide activations: give up co-residency, take what the SM has
    ns = (int)std::min<long long>((G2_SMEM_BUDGET - fixed) / G2_STAGE_BYTES, std::max<long long>(per_cta, 3));
    ns = std::min(ns, G2_MAX_STAGES);
  }
  return std::max(ns, 2);
}

static unsigned long long* g_prof_base = nullptr;
static long long g_prof_cap = 0, g_prof_next = 0;
void gemv2_set_prof(unsigned long long* base, long long n_launches) { g_prof_base = base; g_prof_cap = n_launches; g_prof_next = 0; }
long long prof_count() { return g_prof_next; }
long long prof_next_slot() { return (g_prof_base && g_prof_next < g_prof_cap) ? g_prof_next++ : -1; }
unsigned long long* prof_slot_ptr(long long slot) { return slot < 0 ? nullptr : g_prof_base + (size_t)slot * 148 * 8; }

void gemv2_launch(const GemvParams& p, int mode, float* ws_part, unsigned* tickets, cudaStream_t stream) {
  Gemv2Params pp;
  pp.g = p;
  pp.prof = prof_slot_ptr(prof_next_slot());
  const bool glu = p.act != 0;
  pp.rows8 = g2_rows8(p.N, glu) ? 1 : 0;
  make_weight_tmap(&pp.tmap, p.w, p.N, p.K, (glu || pp.rows8) ? 8 : 16);
  pp.ws_part = ws_part;
  pp.tickets = tickets;
  pp.p_max = gemv2_pmax(p.N, p.K, glu);
  pp.whole_tiles = g2_whole_tiles(p.N, p.K, glu) ? 1 : 0;
  pp.n_stages = g2_pick_stages(p, glu, pp.whole_tiles != 0);
  if (mode == 1) launch_gemv2<false, 1>(pp, stream);
  else if (glu) launch_gemv2<true, 0>(pp, stream);
  else launch_gemv2<false, 0>(pp, stream);
}

}  // namespace nxdi
