// Decode GEMV building blocks shared by the stand-alone kernel (gemv2.cu) and the persistent decode-step kernel
// (decode_step.cu): one TMA producer thread + 8 consumer warps around an mbarrier ring of 16 KB stages.  See gemv2.cu for the
// design notes.  A "phase" is one skinny GEMM  Y[T<=8, N] = f(rmsnorm(X)[T,K] · W[N,K]^T)  (+ fused LL all-reduce).
#pragma once
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
constexpr int G2_SMEM_CORES = 224 * 1024;    // default ring budget.  Measured (profiles/decode_r2.md): one deep ring per SM (224 KB) beats
                                             // two co-resident 108 KB CTAs at TP1 (3.17 vs 3.64 ms/step) and on TP8 shard shapes (1.22 vs 1.36)

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
// Explicit shared-space accesses by 32-bit shared address.  The staging areas are carved out of dynamic shared memory at run
// time; through plain pointers the compiler loses the address space and emits GENERIC loads/stores (LD.E / ST.E) for every
// fragment — measured as a uniform 1.5x slowdown of the streaming loop.
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// element offset (bf16) of 16-byte chunk `chunk` (0..7) of 128-byte line `line` inside a 128B-swizzled stage
__device__ __forceinline__ int swz128(int line, int chunk) { return line * 64 + ((chunk ^ (line & 7)) << 3); }
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// host helpers implemented in gemv2.cu
void make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K, int box_rows, int wt = 0);
int g2_num_sms_public();
int g2_max_inflight();   // NXDI_B200_GEMV_INFLIGHT

// One skinny GEMM as seen by a CTA.  Plain data: lives in kernel parameter space (gemv2.cu) or in global memory (decode_step.cu).
struct G2Phase {
  const CUtensorMap* tmap;   // W as {64 k, N rows, K/64 groups}, box {64, 16 (8 for GLU / rows8), 8}, SWIZZLE_128B
  const void* x;             // [T, K] bf16
  const void* bias;          // [N] bf16 or null
  const void* norm_w;        // [K] bf16 or null (fused RMSNorm of x)
  const void* residual;      // [T, N] bf16 or null
  void* y;                   // [T, N_out] bf16
  float* ws_part;            // [n_tiles * p_max][128] fp32 stream-K partials
  unsigned* tickets;         // [n_tiles]
  unsigned long long* prof;  // debug timeline: [8] u64 per CTA or null
  int T, N, K, ldx, ldy;
  float eps, norm_offset;
  int act;                   // 0 none, 1 silu*up, 2 gelu_tanh*up, 3 gelu*up  (!= 0: GLU tiles)
  int p_max;
  int rows8;                 // 8-row tiles (plain epilogues, narrow N)
  int whole_tiles;           // CTAs own whole tiles (no stream-K fix-up)
  int grid;                  // CTAs that share this phase's units
  int parity, call;          // fused all-reduce: receive-buffer parity and call index (tag) of this collective
  // weight-only 8-bit weights (wt: 0 bf16, 1 int8, 2 fp8-e4m3): the SAME ring streams bytes — a stage is then 16 rows x 1024 k — and
  // the consumers expand them to f16 pairs in registers (exact) for an f16 MMA against x converted to f16 once in the prologue;
  // acc * wscale[n] (per output channel, or [0] per tensor) in the epilogue
  const float* wscale;
  int wscale_n;
  int wt;
};
__host__ __device__ __forceinline__ int g2_kc(int wt) { return wt ? 2 * G2_KC : G2_KC; }   // k elements per 16 KB stage

// Symmetric-workspace constants of a launch (same for every all-reduce phase)
struct G2Symm {
  float* recv[SYMM_MAX_RANKS];
  const uint32_t* step;
  int rank, world, n_max;
};

// The CTA's ring + staging areas; `stage`/`phase bit` cursors live with the caller (they run on across phases).
struct G2Smem {
  uint8_t* stage_base;   // [NS][16 KB], 1024-aligned
  uint8_t* xs;           // [T][2*Kp+64]
  float* red;            // [8][128]
  float* rstd_s;         // [64]
  uint64_t* full_bar;    // [NS]
  uint64_t* empty_bar;   // [NS]
  uint64_t* x_bar;       // activations landed (one bulk copy per token row)
  int* s_flag;
  int NS;
  int max_inflight;      // producer: at most this many stages issued-but-not-landed (0 = no cap).  A deep ring buffers consumer
                         // stalls, but every byte in flight queues ahead of the step's dependent loads (x, flags): see decode_step.cu
};

struct G2Units {   // unit range of CTA `c` in a phase
  int n_chunks, n_tiles, TR;
  long long U, u_beg, u_end;
  bool rows8;
};
__device__ __forceinline__ G2Units g2_units(const G2Phase& p, bool GLU, int c) {
  G2Units u;
  u.n_chunks = (p.K + g2_kc(p.wt) - 1) / g2_kc(p.wt);
  u.rows8 = !GLU && p.rows8 != 0;
  u.TR = u.rows8 ? 8 : 16;
  u.n_tiles = GLU ? ((p.N >> 1) + 7) >> 3 : (p.N + u.TR - 1) / u.TR;
  u.U = (long long)u.n_tiles * u.n_chunks;
  const int G = p.grid;
  if (c >= G) { u.u_beg = u.u_end = 0; return u; }
  u.u_beg = p.whole_tiles ? ((long long)u.n_tiles * c / G) * u.n_chunks : (u.U * c) / G;
  u.u_end = p.whole_tiles ? ((long long)u.n_tiles * (c + 1) / G) * u.n_chunks : (u.U * (c + 1)) / G;
  return u;
}

// ---- producer: ONE thread streams this CTA's units of the phase through the ring ----
// One issuing thread sustains one copy per ~150-230 ns whatever its size (tools/bench_stream.cu: 4 KB boxes 2.7 TB/s,
// 8 KB 5.4 TB/s, 16 KB 7.2 TB/s chip-wide): the stage is therefore ONE 16 KB box (two 8 KB boxes for GLU tiles).
__device__ __forceinline__ void g2_produce(const G2Phase& p, const G2Smem& sm, const int c, int& stage, uint32_t& ph, int& issued) {
  const bool GLU = p.act != 0;
  const G2Units un = g2_units(p, GLU, c);
  const int N = p.N, NS = sm.NS, n_chunks = un.n_chunks, TR = un.TR;
  const bool rows8 = un.rows8;
  uint8_t* stage_base = sm.stage_base;
  uint64_t* full_bar = sm.full_bar;
  uint64_t* empty_bar = sm.empty_bar;
  const CUtensorMap* tmap = p.tmap;
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
  int tile = (int)(un.u_beg / n_chunks), chunk = (int)(un.u_beg % n_chunks);
  const int count = (int)(un.u_end - un.u_beg);
  const int W = sm.max_inflight;
  for (int i = 0; i < count; ++i, ++issued) {
    if (W > 0 && issued >= W) {   // the stage issued W units ago must have landed before another one goes out
      const int st = stage >= W ? stage - W : stage - W + NS;
      mbar_wait(&full_bar[st], stage >= W ? ph : ph ^ 1u);
    }
    mbar_wait(&empty_bar[stage], ph ^ 1u);
    mbar_expect_tx(&full_bar[stage], rows8 ? G2_STAGE_BYTES / 2 : G2_STAGE_BYTES);
    uint8_t* dst = stage_base + (size_t)stage * G2_STAGE_BYTES;
    // k groups past K/64 and rows past N are zero-filled by the TMA unit (and not fetched)
    if (GLU) {  // 8 gate rows, then 8 up rows
      tma_load_3d(dst, tmap, 0, tile * 8, chunk * G2_KG, &full_bar[stage]);
      tma_load_3d(dst + G2_STAGE_BYTES / 2, tmap, 0, (N >> 1) + tile * 8, chunk * G2_KG, &full_bar[stage]);
    } else {
      tma_load_3d(dst, tmap, 0, tile * TR, chunk * G2_KG, &full_bar[stage]);
    }
    if (++chunk == n_chunks) { chunk = 0; ++tile; }
    if (++stage == NS) { stage = 0; ph ^= 1u; }
  }
}

// ---- consumers: the 256 threads of warps 0..7.  `wait_dep()` blocks until the phase's inputs are complete (griddepcontrol.wait
//      in the stand-alone kernel, a device-side counter in the persistent one); everything before it overlaps the producer of x.
// 4 one-byte weights (one 32-bit word, k ascending) -> two f16x2 registers (k, k+1), (k+2, k+3); exact for both formats
template <int WT>
__device__ __forceinline__ void g2_expand4(uint32_t w, uint32_t& lo, uint32_t& hi) {
  if (WT == 2) {
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(lo) : "h"((unsigned short)(w & 0xffffu)));
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(hi) : "h"((unsigned short)(w >> 16)));
  } else {
    // int8: bias to unsigned, splice under the exponent of 1024.0 (f16 ulp 1 in [1024, 2048)), subtract 1024 + 128
    const uint32_t u = w ^ 0x80808080u;
    const uint32_t a = __byte_perm(u, 0x64646464u, 0x5140), b = __byte_perm(u, 0x64646464u, 0x5342);
    const uint32_t k1152 = 0x64806480u;
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(k1152));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(k1152));
  }
}
// two fp32 -> f16x2 with saturation to the finite range (activations of the 8-bit weight paths)
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <bool GLU, int MODE, int WT = 0, class WaitDep>
__device__ __forceinline__ void g2_consume(const G2Phase& p, const G2Symm& sy, const G2Smem& sm, const int c, const int tid,
                                           int& stage, uint32_t& lap, uint32_t& xph, WaitDep&& wait_dep) {
  const int lane = tid & 31, warp = tid >> 5;
  const int K = p.K, N = p.N, T = p.T, NS = sm.NS, G = p.grid;
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* BIAS = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  const __nv_bfloat16* RES = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);
  const G2Units un = g2_units(p, GLU, c);
  const int n_chunks = un.n_chunks, n_tiles = un.n_tiles, TR = un.TR;
  const bool rows8 = un.rows8;
  const long long U = un.U, u_beg = un.u_beg, u_end = un.u_end;
  constexpr int KC = WT ? 2 * G2_KC : G2_KC;   // k elements per stage
  const int Kp = n_chunks * KC;
  // row stride of x in shared memory: the 16-byte offset (8-bit paths: lanes of a quad are 32 bytes apart) or the 64-byte
  // offset (bf16: 16 bytes apart) keeps the 8-lane LDS.128 phases of two token rows in disjoint banks
  const int xs_stride = Kp * 2 + (WT != 0 ? 16 : 64);
  const uint32_t stage_u = smem_u32(sm.stage_base), xs_u = smem_u32(sm.xs), red_u = smem_u32(sm.red), rstd_u = smem_u32(sm.rstd_s);
  uint64_t* full_bar = sm.full_bar;
  uint64_t* empty_bar = sm.empty_bar;
  int& s_flag = *sm.s_flag;
  const int g = lane >> 2, t4 = lane & 3;
  const int ctid = tid;  // 0..255
  unsigned long long* prof = p.prof ? p.prof + (size_t)c * 8 : nullptr;
  if (prof && ctid == 0) { prof[0] = gtimer(); prof[1] = clock64(); }
  const int nvec = K >> 3;
  const bool has_norm = p.norm_w != nullptr;
  // ---- before the dependency resolves: everything that does not depend on the previous kernel ----
  // gamma -> registers (this thread's vectors ctid, ctid+256, ...); zero the k padding of every x row (K..Kp)
  constexpr int G2_MAXV = 7;   // K <= 14336
  uint4 gq[G2_MAXV];
  const int per_thread = (nvec - ctid + 255) / 256;   // vectors of this thread
  if (has_norm) {
    const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
#pragma unroll
    for (int j = 0; j < G2_MAXV; ++j)
      if (j < per_thread) gq[j] = ldg_cached(gw + ctid + 256 * j);
  }
  if (Kp != K) {
    const int padv = (Kp - K) >> 3;
    for (int i = ctid; i < T * padv; i += 256)
      sts128(xs_u + (i / padv) * xs_stride + (nvec + i % padv) * 16, make_uint4(0u, 0u, 0u, 0u));
  }
  wait_dep();
  if (prof && ctid == 0) prof[2] = clock64();
  // all-reduce tag of this call: (device step counter << 8 | call index) + 1; the counter is bumped by the host-enqueued
  // begin_step op before the first collective of every forward (parallel/symm.py), so graph replays get fresh tags
  uint32_t ar_tag = 1u;
  if (MODE == 1) ar_tag = ll_tag(sy.step, p.call);
  // ---- X prologue: ONE bulk copy (TMA, cp.async.bulk) per token row brings the activations into shared memory — a single
  //      request stream per CTA instead of 1024 scattered 16-byte loads (measured: 1-3.5 us -> see profiles/decode_r2.md).
  //      With a fused RMSNorm the rows are then scaled by gamma IN PLACE (bf16); the per-token 1/rms is a scalar, so it is
  //      applied to the fp32 accumulators in the epilogue instead of to x.
  if (ctid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy writes to xs (previous phase) vs the bulk copy
    mbar_expect_tx(sm.x_bar, (uint32_t)(T * K * 2));
    for (int t = 0; t < T; ++t)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(xs_u + t * xs_stride),
                   "l"(X + (size_t)t * p.ldx), "r"(K * 2), "r"(smem_u32(sm.x_bar))
                   : "memory");
  }
  mbar_wait(sm.x_bar, xph);
  xph ^= 1u;
  if (has_norm || WT != 0) {
    // one in-place pass over x: sum of squares (fused RMSNorm), x gamma, and for the 8-bit weight paths bf16 -> f16
    const float o = p.norm_offset;
#pragma unroll
    for (int t = 0; t < GEMV_MAX_T; ++t) {
      if (t < T) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < G2_MAXV; ++j) {
          if (j < per_thread) {
            const uint32_t a = xs_u + t * xs_stride + (ctid + 256 * j) * 16;
            uint4 w = lds128(a);
            float f[8] = {bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y), bf16lo(w.z), bf16hi(w.z), bf16lo(w.w), bf16hi(w.w)};
            if (has_norm) {
              const uint4 gm = gq[j];
              const float gf[8] = {bf16lo(gm.x), bf16hi(gm.x), bf16lo(gm.y), bf16hi(gm.y), bf16lo(gm.z), bf16hi(gm.z), bf16lo(gm.w), bf16hi(gm.w)};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                acc += f[e] * f[e];
                f[e] *= gf[e] + o;
              }
            }
            if (WT != 0) {
              w = make_uint4(pack_f16_sat(f[0], f[1]), pack_f16_sat(f[2], f[3]), pack_f16_sat(f[4], f[5]), pack_f16_sat(f[6], f[7]));
            } else {
              w = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
            sts128(a, w);
          }
        }
        if (has_norm) {
          acc = warp_sum(acc);
          if (lane == 0) sts_f32(rstd_u + (warp * 8 + t) * 4, acc);
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
    for (int w = 0; w < G2_CONSUMER_WARPS; ++w) tot += lds_f32(rstd_u + (w * 8 + col) * 4);
    return rsqrtf(tot / (float)K + p.eps);
  };

  const bool tok_ok = g < T;
  const uint32_t xrow = xs_u + g * xs_stride;
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};

  // epilogue operands of the tile being streamed (bias / residual of this thread's (col, row)): fetched when the tile STARTS,
  // so their L2 latency hides behind the weight stream instead of sitting on the critical path after the last stage
  float pre_b0 = 0.f, pre_b1 = 0.f, pre_r = 0.f, pre_s0 = 1.f, pre_s1 = 1.f;
  auto prefetch_epilogue = [&](int tile) {
    pre_b0 = pre_b1 = pre_r = 0.f;
    pre_s0 = pre_s1 = 1.f;
    if (ctid >= 128) return;
    const int col = ctid >> 4, row = ctid & 15;
    if (col >= T) return;
    if (GLU) {
      const int n = tile * 8 + row, half = N >> 1;
      if (row < 8 && n < half) {
        if (BIAS != nullptr) {
          pre_b0 = __bfloat162float(BIAS[n]);
          pre_b1 = __bfloat162float(BIAS[half + n]);
        }
        if (WT != 0) {
          pre_s0 = __ldg(p.wscale + (p.wscale_n == 1 ? 0 : n));
          pre_s1 = __ldg(p.wscale + (p.wscale_n == 1 ? 0 : half + n));
        }
      }
    } else {
      const int n = tile * TR + row;
      if (row < TR && n < N) {
        if (WT != 0) pre_s0 = __ldg(p.wscale + (p.wscale_n == 1 ? 0 : n));
        if (MODE == 0) {
          if (BIAS != nullptr) pre_b0 = __bfloat162float(BIAS[n]);
          if (RES != nullptr) pre_r = ldg_act_bf16(RES + (size_t)col * p.ldy + n);
        }
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
      const float gate = v_gate_or_val * (rs * pre_s0) + pre_b0, up = v_up * (rs * pre_s1) + pre_b1;
      const float a = p.act == 1 ? silu(gate) : (p.act == 2 ? gelu_tanh(gate) : gelu_erf(gate));
      Y[(size_t)col * p.ldy + n] = __float2bfloat16(a * up);
    } else {
      if (row >= TR) return;
      const int n = tile * TR + row;
      if (n >= N) return;
      float v = v_gate_or_val * (rstd_of(col) * pre_s0);
      if (MODE == 0) {
        Y[(size_t)col * p.ldy + n] = __float2bfloat16(v + pre_b0 + pre_r);
      } else {
        // LL all-reduce phase 1: {value, tag} straight into every peer's slot (including mine)
        const G2Symm& s = sy;
        const size_t off = (((size_t)(p.parity * s.world + s.rank) * 8 + col) * s.n_max + n) * 2;
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
  while (seg_beg < count) {
    const int seg_len = min(n_chunks - chunk_first, count - seg_beg);
    const long long u = u_beg + seg_beg + seg_len - 1;  // last unit of this segment (global)
    prefetch_epilogue(cur_tile);
    for (int i = 0; i < seg_len; ++i) {
      const int chunk = chunk_first + i;
      mbar_wait(&full_bar[stage], lap);
      const uint32_t sA = stage_u + stage * G2_STAGE_BYTES;
      // stage layout: line = kg * rows + row (rows = 16; GLU / rows8: two halves of [8 kg][8 rows])
      if (WT != 0) {
        // 8-bit weights: a 128-byte line holds 128 k of one row; this thread takes 16-byte chunks t4 and t4 + 4 (16 k each) of
        // rows g and g + 8, expands them to f16 pairs and multiplies them with the matching 32 bytes of x (f16): 4 MMAs per chunk.
        // Any k -> MMA-slot assignment is fine as long as A and B agree: pair m of a chunk (k = 2m, 2m + 1) is slot pair m.
        const uint32_t xk8 = xrow + (chunk * KC + warp * 128 + t4 * 16) * 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ch = (j << 2) + t4;
          uint4 q0, q1 = make_uint4(0u, 0u, 0u, 0u);
          if (GLU) {
            q0 = lds128(sA + swz128(warp * 8 + g, ch) * 2);
            q1 = lds128(sA + G2_STAGE_BYTES / 2 + swz128(warp * 8 + g, ch) * 2);
          } else if (rows8) {
            q0 = lds128(sA + swz128(warp * 8 + g, ch) * 2);
          } else {
            q0 = lds128(sA + swz128(warp * 16 + g, ch) * 2);
            q1 = lds128(sA + swz128(warp * 16 + g + 8, ch) * 2);
          }
          uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
          if (tok_ok) {
            x0 = lds128(xk8 + j * 128);
            x1 = lds128(xk8 + j * 128 + 16);
          }
          uint32_t pa[8], pb[8];
          g2_expand4<WT>(q0.x, pa[0], pa[1]); g2_expand4<WT>(q0.y, pa[2], pa[3]);
          g2_expand4<WT>(q0.z, pa[4], pa[5]); g2_expand4<WT>(q0.w, pa[6], pa[7]);
          g2_expand4<WT>(q1.x, pb[0], pb[1]); g2_expand4<WT>(q1.y, pb[2], pb[3]);
          g2_expand4<WT>(q1.z, pb[4], pb[5]); g2_expand4<WT>(q1.w, pb[6], pb[7]);
          const uint32_t xb[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const uint32_t a[4] = {pa[2 * m], pb[2 * m], pa[2 * m + 1], pb[2 * m + 1]};
            const uint32_t b[2] = {xb[2 * m], xb[2 * m + 1]};
            if (m & 1) mma_f16_16816(c1, a, b);
            else mma_f16_16816(c0, a, b);
          }
        }
      }
      const uint32_t xk = xrow + (chunk * G2_KC + warp * 64 + t4 * 8) * 2;
#pragma unroll
      for (int j = 0; j < (WT != 0 ? 0 : 2); ++j) {
        const int ch = (j << 2) + t4;
        uint4 a0, a1 = make_uint4(0u, 0u, 0u, 0u);
        if (GLU) {
          a0 = lds128(sA + swz128(warp * 8 + g, ch) * 2);
          a1 = lds128(sA + G2_STAGE_BYTES / 2 + swz128(warp * 8 + g, ch) * 2);
        } else if (rows8) {
          a0 = lds128(sA + swz128(warp * 8 + g, ch) * 2);
        } else {
          a0 = lds128(sA + swz128(warp * 16 + g, ch) * 2);
          a1 = lds128(sA + swz128(warp * 16 + g + 8, ch) * 2);
        }
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (tok_ok) xv = lds128(xk + j * 64);
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
    const uint32_t r = red_u + warp * 512;
    sts_f32(r + (g * 8 + 2 * t4) * 4, c0[0] + c1[0]);
    sts_f32(r + (g * 8 + 2 * t4 + 1) * 4, c0[1] + c1[1]);
    sts_f32(r + ((g + 8) * 8 + 2 * t4) * 4, c0[2] + c1[2]);
    sts_f32(r + ((g + 8) * 8 + 2 * t4 + 1) * 4, c0[3] + c1[3]);
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
        va += lds_f32(red_u + (w * 128 + row * 8 + col) * 4);
        if (GLU) vb += lds_f32(red_u + (w * 128 + ((row + 8) & 15) * 8 + col) * 4);
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
      float* my = p.ws_part + ((size_t)cur_tile * p.p_max + slot) * 128;
      if (ctid < 128) {
        const int col = ctid >> 4, row = ctid & 15;
        my[row * 8 + col] = va;  // natural (row, col) layout: rows 0-7 gate / 8-15 up for GLU tiles
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ctid == 0) s_flag = (atomicAdd(&p.tickets[cur_tile], 1u) == (unsigned)(n_parts - 1)) ? 1 : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_flag) {
        __threadfence();
        if (ctid < 128) {
          const int col = ctid >> 4, row = ctid & 15;
          float sa = 0.f, sb = 0.f;
          for (int q = 0; q < n_parts; ++q) {
            const float* pq = p.ws_part + ((size_t)cur_tile * p.p_max + q) * 128;
            sa += __ldcg(pq + row * 8 + col);
            if (GLU) sb += __ldcg(pq + ((row + 8) & 15) * 8 + col);
          }
          finalize(cur_tile, sa, sb);
        }
        if (ctid == 0) p.tickets[cur_tile] = 0;
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

  if (MODE == 1 && c < G) {
    // ---- LL all-reduce phase 2: every CTA polls a strided share of ALL (column, token) slots; the `world` sources of an
    //      element are polled together (independent loads in flight: one L2 round trip per poll, not `world`) ----
    const G2Symm& s = sy;
    float* my_recv = s.recv[0];
#pragma unroll
    for (int d = 1; d < SYMM_MAX_RANKS; ++d)
      if (d == s.rank) my_recv = s.recv[d];
    int q = 0;
    for (int e = c * 256 + ctid; e < ar_total; e += G * 256, ++q) {
      const int col = e / N, n = e % N;
      const float* slot0 = my_recv + (((size_t)(p.parity * s.world) * 8 + col) * s.n_max + n) * 2;
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

}  // namespace nxdi
