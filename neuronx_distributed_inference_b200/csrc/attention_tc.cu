// Flash attention (prefill) on the 5th-generation tensor cores: S = Q K^T and O += P V are tcgen05.mma instructions with their
// accumulators in TENSOR MEMORY; Q / K / V tiles arrive through TMA (cp.async.bulk.tensor, SWIZZLE_128B); the softmax owns one
// query row per thread (tcgen05.ld of its TMEM lane), so no cross-thread reductions exist anywhere.
//
//   CTA = one (batch, q head, 128-query tile).  192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
//   warps 2..5 = softmax / correction / epilogue (TMEM lane quarter = warp % 4).
//   per 128-key tile j:   S_j = Q K_j^T          8 x UMMA 128x128x16 (K-major A = Q, K-major B = K_j)        -> TMEM cols [0,128)
//                         P_j = exp2(S_j*c - m)  registers -> bf16 -> shared memory (128B-swizzled, K-major A operand)
//                         O  += P_j V_j          8 x UMMA 128x128x16 (B = V_j in its natural [key][d] layout = MN-major) -> TMEM cols [128,256)
//   The running maximum is LAZY (FlashAttention-4 style): O and l are rescaled only when a row's maximum grows by more than 2^8,
//   decided per warp, so the TMEM read-modify-write of O happens on a handful of tiles instead of all of them.
//   S_{j+1} is issued right after P_j V_j, so the tensor pipe works on the next scores while the softmax of tile j+1 ... runs on
//   the CUDA cores (single S / P buffers: the pipeline is MMA -> softmax -> MMA with the QK^T of the next tile overlapped).
// Masks are arguments: causal, sliding window, bidirectional (causal = 0), per-head sink logit, tanh soft-cap.
// reference kernel contract: attention_cte / flash_fwd (modules/attention/attention_base.py:603-630,722-744).
#include <cfloat>
#include <stdexcept>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 128;
constexpr int FA_THREADS = 192;
constexpr int FA_KV_STAGES = 2;
constexpr int FA_TILE_BYTES = 128 * 128 * 2;   // one [128 rows][128 cols] bf16 tile = two 16 KB halves of 64 columns
constexpr float FA_LOG2E = 1.4426950408889634f;

struct FaParams {
  CUtensorMap tm_q, tm_k, tm_v;   // 2-D views [B*T rows, H*D cols]; box {64, 128}, SWIZZLE_128B
  __nv_bfloat16* out;             // [B, T, Hq, D]
  const float* sinks;             // [Hq] or null
  int B, T, Hq, Hkv, causal, window;
  float scale_log2, softcap;
};

__device__ __forceinline__ uint32_t fa_s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void fa_mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fa_s32(b)), "r"(n)); }
__device__ __forceinline__ void fa_mb_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fa_s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fa_mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fa_s32(b)) : "memory"); }
__device__ __forceinline__ void fa_mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(fa_s32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void fa_tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(fa_s32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(fa_s32(bar)) : "memory");
}
// K-major operand tile [rows][64 cols] of 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t fa_desc_kmajor(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((fa_s32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (V in its natural [key][d] layout): canonical SW128 layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in 16-byte
// units — a 128-byte row holds 64 contiguous MN (= d) elements of ONE k (= key); 8 keys form a 1024-byte group (SBO);
// the second 64-column half of the tile sits LBO = 16 KB further.
__device__ __forceinline__ uint64_t fa_desc_mnmajor(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((fa_s32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)((FA_TILE_BYTES / 2) >> 4) << 16;   // leading byte offset: next 64-element block along MN
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row (key) group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, M = 128, N = 128; bit 16 = B is MN-major
__host__ __device__ constexpr uint32_t fa_idesc(bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void fa_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fa_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fa_s32(bar)) : "memory");
}
__device__ __forceinline__ void fa_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void fa_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

__global__ void __launch_bounds__(FA_THREADS, 1) attention_prefill_tc_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (fa_s32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                    // [2 d-halves][128 q][64 d]
  uint8_t* sK = sQ + FA_TILE_BYTES;                      // [stages][2 d-halves][128 keys][64 d]
  uint8_t* sV = sK + FA_KV_STAGES * FA_TILE_BYTES;       // [stages][2 d-halves][128 keys][64 d]
  uint8_t* sP = sV + FA_KV_STAGES * FA_TILE_BYTES;       // [2 key-halves][128 q][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_TILE_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // [stages]  K and V of a stage landed
  uint64_t* kv_empty = kv_full + FA_KV_STAGES;   // [stages]  P V of the stage retired
  uint64_t* s_full = kv_empty + FA_KV_STAGES;    // S_j complete
  uint64_t* p_full = s_full + 1;           // P_j written (and O rescaled): 4 softmax warps arrive
  uint64_t* pv_done = p_full + 1;          // P_j V_j retired: P buffer and O accumulator are quiescent
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.T + FA_BM - 1) / FA_BM;
  const int qt = n_qt - 1 - (int)(blockIdx.x % n_qt);      // heavy (late) query tiles first
  const int bh = blockIdx.x / n_qt;
  const int b = bh / p.Hq, h = bh % p.Hq, hk = h / (p.Hq / p.Hkv);
  const int q0 = qt * FA_BM;
  // key tiles this query tile attends to
  int j_hi = p.causal ? qt : (p.T + FA_BN - 1) / FA_BN - 1;
  j_hi = min(j_hi, (p.T + FA_BN - 1) / FA_BN - 1);
  int j_lo = 0;
  if (p.window > 0) j_lo = max(0, q0 - p.window + 1) / FA_BN;
  const int n_tiles = j_hi - j_lo + 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_v) : "memory");
    fa_mb_init(q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      fa_mb_init(&kv_full[s], 1);
      fa_mb_init(&kv_empty[s], 1);
    }
    fa_mb_init(s_full, 1);
    fa_mb_init(p_full, 4);
    fa_mb_init(pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fa_s32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int row_q = b * p.T + q0, colq = h * FA_D, colk = hk * FA_D;
      fa_mb_expect(q_full, FA_TILE_BYTES);
      fa_tma_2d(sQ, &p.tm_q, colq, row_q, q_full);
      fa_tma_2d(sQ + FA_TILE_BYTES / 2, &p.tm_q, colq + 64, row_q, q_full);
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % FA_KV_STAGES;
        const uint32_t ph = (uint32_t)((i / FA_KV_STAGES) & 1);
        fa_mb_wait(&kv_empty[s], ph ^ 1u);
        fa_mb_expect(&kv_full[s], 2 * FA_TILE_BYTES);
        const int row_k = b * p.T + (j_lo + i) * FA_BN;
        fa_tma_2d(sK + s * FA_TILE_BYTES, &p.tm_k, colk, row_k, &kv_full[s]);
        fa_tma_2d(sK + s * FA_TILE_BYTES + FA_TILE_BYTES / 2, &p.tm_k, colk + 64, row_k, &kv_full[s]);
        fa_tma_2d(sV + s * FA_TILE_BYTES, &p.tm_v, colk, row_k, &kv_full[s]);
        fa_tma_2d(sV + s * FA_TILE_BYTES + FA_TILE_BYTES / 2, &p.tm_v, colk + 64, row_k, &kv_full[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = fa_idesc(false), idesc_pv = fa_idesc(true);
      fa_mb_wait(q_full, 0);
      auto issue_s = [&](int i) {   // S_i = Q K_i^T
        const int s = i % FA_KV_STAGES;
        fa_mb_wait(&kv_full[s], (uint32_t)((i / FA_KV_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) {
          // 16 d = 32 bytes inside the 128-byte swizzle atom; the second 64-d half is a separate 16 KB block
          const uint64_t da = fa_desc_kmajor(sQ + (k >> 2) * (FA_TILE_BYTES / 2)) + (uint64_t)((k & 3) * 2);
          const uint64_t db = fa_desc_kmajor(sK + s * FA_TILE_BYTES + (k >> 2) * (FA_TILE_BYTES / 2)) + (uint64_t)((k & 3) * 2);
          fa_umma(tmem_S, da, db, idesc_qk, k != 0 ? 1u : 0u);
        }
        fa_commit(s_full);
      };
      issue_s(0);
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % FA_KV_STAGES;
        fa_mb_wait(p_full, (uint32_t)(i & 1));     // P_i in shared memory, O rescaled, S_i fully read
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          // A = P: 16 keys = 32 bytes inside the atom, second 64-key half 16 KB further.
          // B = V (MN-major): 16 keys = 16 rows of 128 bytes = 2048 bytes further in BOTH d halves (LBO covers the halves)
          const uint64_t da = fa_desc_kmajor(sP + (k >> 2) * (FA_TILE_BYTES / 2)) + (uint64_t)((k & 3) * 2);
          const uint64_t db = fa_desc_mnmajor(sV + s * FA_TILE_BYTES + k * 2048);
          fa_umma(tmem_O, da, db, idesc_pv, (i | k) != 0 ? 1u : 0u);
        }
        fa_commit(&kv_empty[s]);      // K_i / V_i stage reusable once these retire
        fa_commit(pv_done);           // P buffer / O accumulator quiescent
        if (i + 1 < n_tiles) issue_s(i + 1);
      }
    }
  } else {
    // ================= softmax / correction / epilogue: one query row per thread =================
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // row inside the tile = TMEM lane
    const int qi = q0 + r;                        // query index inside the sequence
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int i = 0; i < n_tiles; ++i) {
      const int key0 = (j_lo + i) * FA_BN;
      fa_mb_wait(s_full, (uint32_t)(i & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- the whole S row of this thread -> registers (ONE TMEM read; 128 fp32)
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) fa_tmem_ld32(tmem_S + lane_sel + 32 * c, v[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float scl = sc;
      if (p.softcap > 0.f) {   // tanh soft-cap: scores move to the log2 domain here, the scale below becomes 1
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            v[c][e] = __float_as_uint(p.softcap * tanhf(__uint_as_float(v[c][e]) * (sc / FA_LOG2E) / p.softcap) * FA_LOG2E);
        scl = 1.f;
      }
      // masks only where a tile can contain invisible keys (sequence end, causal diagonal, window edge): CTA-uniform test
      const bool need_mask = (key0 + FA_BN > p.T) || (p.causal && key0 + FA_BN - 1 > q0) || (p.window > 0 && key0 <= q0 + FA_BM - 1 - p.window);
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int kj = key0 + 32 * c + e;
            const bool ok = kj < p.T && (!p.causal || kj <= qi) && (p.window <= 0 || kj > qi - p.window);
            if (!ok) v[c][e] = 0xff800000u;   // -inf
          }
      }
      // row maximum in the RAW domain (the scale is positive), four independent chains
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 32; ++e) mx4[e & 3] = fmaxf(mx4[e & 3], __uint_as_float(v[c][e]));
      const float mx_raw = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float mx = (mx_raw == -INFINITY) ? -INFINITY : mx_raw * scl;
      // lazy running maximum: move it only when this tile exceeds it by more than 2^8 (decided per warp so that the TMEM
      // read-modify-write below stays warp-uniform); rows that keep their maximum use alpha = 1
      float m_new = m_run;
      const bool grow = mx > m_run + 8.f || (m_run == -INFINITY && mx > -INFINITY);
      if (grow) m_new = mx;
      const bool any = __any_sync(0xffffffffu, grow);
      const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      // the previous P V must have retired before P is overwritten / O is touched
      if (i > 0) fa_mb_wait(pv_done, (uint32_t)((i - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (any && i > 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < FA_D; c0 += 32) {
          uint32_t o[32];
          fa_tmem_ld32(tmem_O + lane_sel + c0, o);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          fa_tmem_st32(tmem_O + lane_sel + c0, o);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      l_run *= (grow ? alpha : 1.f);
      m_run = m_new;
      // ---- P = exp2(s * scale - m) -> bf16 -> shared memory (row r of both 64-key halves, 16-byte chunks XOR-swizzled by r & 7)
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(v[c][e]), scl, neg_m));        // -inf -> 0
          const float p1 = exp2f(fmaf(__uint_as_float(v[c][e + 1]), scl, neg_m));
          // the sum uses the bf16-rounded probabilities, exactly what the tensor core multiplies
          const uint32_t w = pack_bf16(p0, p1);
          ls4[(e >> 1) & 3] += bf16lo(w) + bf16hi(w);
          pk[e >> 1] = w;
        }
        const int c0 = 32 * c;
        uint8_t* rowp = sP + (c0 >> 6) * (FA_TILE_BYTES / 2) + r * 128;
        const int ch0 = (c0 & 63) >> 3;   // first 16-byte chunk of this 32-key group inside the 128-byte row
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          *reinterpret_cast<uint4*>(rowp + (((ch0 + cc) ^ (r & 7)) << 4)) = make_uint4(pk[4 * cc], pk[4 * cc + 1], pk[4 * cc + 2], pk[4 * cc + 3]);
      }
      const float lsum = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      l_run += lsum;
      // generic-proxy writes of P must be visible to the tensor core (async proxy); S reads are complete
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fa_mb_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global
    fa_mb_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float l = l_run;
    if (p.sinks != nullptr) {
      const float sk = p.sinks[h] * FA_LOG2E;
      l = (m_run == -INFINITY) ? 1.f : l + exp2f(sk - m_run);
    }
    const float inv = (m_run == -INFINITY || l == 0.f) ? 0.f : 1.f / l;
    __nv_bfloat16* dst = p.out + (((size_t)b * p.T + qi) * p.Hq + h) * FA_D;
#pragma unroll 1
    for (int c0 = 0; c0 < FA_D; c0 += 32) {
      uint32_t v[32];
      fa_tmem_ld32(tmem_O + lane_sel + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (qi < p.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          *reinterpret_cast<uint4*>(dst + c0 + e) =
              make_uint4(pack_bf16(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv),
                         pack_bf16(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv),
                         pack_bf16(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv),
                         pack_bf16(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

typedef CUresult (*FaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void fa_make_map(CUtensorMap* tm, const void* ptr, long long rows, int cols) {
  static FaEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) != cudaSuccess || q == nullptr)
      throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<FaEncodeFn>(q);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (attention) failed: " + std::to_string((int)r));
}

// q [B,T,Hq,128], k / v [B,T,Hkv,128] bf16 contiguous -> out [B,T,Hq,128]
void attention_prefill_tc_launch(const void* q, const void* k, const void* v, void* out, const float* sinks, int B, int T, int Hq, int Hkv,
                                 float scale, int causal, int window, float softcap, cudaStream_t stream) {
  FaParams p{};
  fa_make_map(&p.tm_q, q, (long long)B * T, Hq * FA_D);
  fa_make_map(&p.tm_k, k, (long long)B * T, Hkv * FA_D);
  fa_make_map(&p.tm_v, v, (long long)B * T, Hkv * FA_D);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.sinks = sinks;
  p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.causal = causal; p.window = window;
  p.scale_log2 = scale * FA_LOG2E;
  p.softcap = softcap;
  const size_t smem = (size_t)(2 + 2 * FA_KV_STAGES) * FA_TILE_BYTES + 256 + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(attention_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int n_qt = (T + FA_BM - 1) / FA_BM;
  launch_pdl(attention_prefill_tc_kernel, dim3(B * Hq * n_qt), dim3(FA_THREADS), smem, stream, p);
}

}  // namespace nxdi
