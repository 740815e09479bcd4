// GEMV v2: bulk-async (TMA 1-D, UBLKCP) mbarrier-pipelined, stream-K balanced weight-streaming skinny GEMM
//          Y[T<=8, N] = f( rmsnorm(X)[T,K] · W[N,K]^T )      (+ fused one-shot all-reduce over NVLink)
//
// Why a second design (v1 = gemv.cuh, register-staged LDG): ncu on v1 showed warps >90 % stalled on
// long-scoreboard with DRAM at 51-64 % — in-flight bytes were bounded by registers (2 stages x 8 LDG.128 per
// warp) and a 16-row tile granularity left 15-35 % of the chip idle in the last wave.  v2 fixes both:
//   * ONE producer warp streams W with cp.async.bulk (global -> shared, completion on an mbarrier): a stage is
//     16 weight rows x 256 k (8 KB); up to 20 stages (160 KB) are in flight per SM, independent of registers;
//   * 8 consumer warps wait on the stage's full-barrier, take their 32-k slice as mma.sync A fragments straight
//     from shared memory (row pitch 512+64 B keeps the 8-lane LDS.128 phases bank-conflict free), multiply with
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

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int G2_CONSUMER_WARPS = 8;
constexpr int G2_THREADS = (G2_CONSUMER_WARPS + 1) * 32;  // + producer warp
constexpr int G2_KC = 256;                                 // k elements per stage
constexpr int G2_ROW_PITCH = G2_KC * 2 + 64;               // bytes
constexpr int G2_STAGE_BYTES = 16 * G2_ROW_PITCH;          // 9216
constexpr int G2_MAX_STAGES = 20;

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
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
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
  GemvParams g;
  float* ws_part;     // [n_tiles * p_max][128] fp32 stream-K partials
  unsigned* tickets;  // [n_tiles]
  int p_max;
  int n_stages;
};

template <bool GLU, int MODE>
__global__ void __launch_bounds__(G2_THREADS, 1) gemv2_kernel(const Gemv2Params pp) {
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

  // shared memory carve-up: [stages][16][pitch] | xs[T][2K+64] | red[8][128] f32 | rstd[64] f32 | barriers
  uint8_t* stage_base = smem_raw;
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
  const long long u_beg = (U * c) / G, u_end = (U * (c + 1)) / G;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], G2_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == G2_CONSUMER_WARPS) {
    // =========================== producer warp: stream W with bulk async copies ===========================
    pdl_launch_dependents();
    const int r = lane & 15;
    long long i = 0;
    for (long long u = u_beg; u < u_end; ++u, ++i) {
      const int tile = (int)(u / n_chunks), chunk = (int)(u % n_chunks);
      const int stage = (int)(i % NS);
      const uint32_t ph = (uint32_t)((i / NS) & 1);
      mbar_wait(&empty_bar[stage], ph ^ 1u);
      if (lane == 0) mbar_expect_tx(&full_bar[stage], 16 * G2_KC * 2);
      __syncwarp();
      if (lane < 16) {
        int row;
        if (GLU) {
          row = r < 8 ? min(tile * 8 + r, (N >> 1) - 1) : min((N >> 1) + tile * 8 + (r - 8), N - 1);
        } else {
          row = min(tile * 16 + r, N - 1);
        }
        bulk_g2s(stage_base + (size_t)stage * G2_STAGE_BYTES + r * G2_ROW_PITCH, W + (size_t)row * K + chunk * G2_KC,
                 G2_KC * 2, &full_bar[stage]);
      }
    }
    return;
  }

  // ================================= consumer warps =================================
  const int g = lane >> 2, t4 = lane & 3;
  const int ctid = tid;  // 0..255
  pdl_wait();
  // ---- X prologue: (optional RMSNorm) -> bf16 in shared memory ----
  {
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
      for (int v = ctid; v < nvec; v += 256) {
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
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int t = 0; t < T; ++t) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < G2_CONSUMER_WARPS; ++w) tot += rstd_s[w * 8 + t];
        const float rstd = rsqrtf(tot / (float)K + p.eps);
        uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)t * xs_stride);
        const uint4* gw = reinterpret_cast<const uint4*>(p.norm_w);
        for (int v = ctid; v < nvec; v += 256) {
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
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

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

  long long i = 0;
  int cur_tile = (int)(u_beg / n_chunks);
  long long tile_u0 = u_beg;  // first unit of the current tile handled by this CTA
  for (long long u = u_beg; u < u_end; ++u, ++i) {
    const int chunk = (int)(u % n_chunks);
    const int stage = (int)(i % NS);
    const uint32_t ph = (uint32_t)((i / NS) & 1);
    mbar_wait(&full_bar[stage], ph);
    const uint8_t* sA = stage_base + (size_t)stage * G2_STAGE_BYTES + warp * 64 + t4 * 16;
    const uint4 a0 = *reinterpret_cast<const uint4*>(sA + g * G2_ROW_PITCH);
    const uint4 a1 = *reinterpret_cast<const uint4*>(sA + (g + 8) * G2_ROW_PITCH);
    uint4 xv = make_uint4(0u, 0u, 0u, 0u);
    if (tok_ok) xv = *reinterpret_cast<const uint4*>(xrow + (size_t)(chunk * G2_KC + warp * 32 + t4 * 8) * 2);
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
    // the warp-collective MMA has consumed every lane's fragments: the stage can be refilled
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    const bool tile_done = (chunk == n_chunks - 1) || (u + 1 == u_end);
    if (!tile_done) continue;

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
    cur_tile += 1;
    tile_u0 = u + 1;
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
  return (size_t)T * (K * 2 + 64) + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + 2 * G2_MAX_STAGES * sizeof(uint64_t) + 128;
}

bool gemv2_supported(int T, int K) {
  return K % G2_KC == 0 && g2_fixed_smem(T, K) + 4 * G2_STAGE_BYTES <= 227 * 1024;
}

int gemv2_grid(int N, int K, bool glu) {
  const int n_tiles = glu ? ((N / 2) + 7) / 8 : (N + 15) / 16;
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
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured = true;
  }
  const size_t smem = g2_fixed_smem(pp.g.T, pp.g.K) + (size_t)pp.n_stages * G2_STAGE_BYTES;
  const int grid = gemv2_grid(pp.g.N, pp.g.K, GLU);
  launch_pdl(kern, dim3(grid), dim3(G2_THREADS), smem, stream, pp);
}

void gemv2_launch(const GemvParams& p, int mode, float* ws_part, unsigned* tickets, cudaStream_t stream) {
  Gemv2Params pp;
  pp.g = p;
  pp.ws_part = ws_part;
  pp.tickets = tickets;
  const bool glu = p.act != 0;
  pp.p_max = gemv2_pmax(p.N, p.K, glu);
  const size_t fixed = g2_fixed_smem(p.T, p.K);
  pp.n_stages = (int)std::min<size_t>(G2_MAX_STAGES, (227 * 1024 - fixed) / G2_STAGE_BYTES);
  if (mode == 1) launch_gemv2<false, 1>(pp, stream);
  else if (glu) launch_gemv2<true, 0>(pp, stream);
  else launch_gemv2<false, 0>(pp, stream);
}

}  // namespace nxdi
