// Prefill / large-T GEMM on the 5th-generation tensor cores:  C[M,N] = epilogue( A[M,K] · B[N,K]^T )   (bf16 in, fp32 acc)
//
//   A = activations [tokens, K] (K-major), B = nn.Linear weight [out, K] (K-major)  ->  no transposes anywhere.
//   PERSISTENT: one CTA per SM walks the tile list (M fastest, so the CTAs resident at any moment share a few weight tiles and
//   the weights leave HBM once).  CTA tile = (TM x 128) rows x BN columns:
//     <TM 2, BN 128>  prefill of <= 256 tokens: ONE row of tiles, every weight byte is fetched exactly once (memory bound);
//     <TM 1, BN 256>  larger M: UMMA 128x256x16 reads 12 KB of shared memory per 128 tensor cycles (96 B/clk) — two 128x128
//                     instructions would read 16 KB (128 B/clk = the whole smem bandwidth; measured 72 % of cuBLAS, stuck);
//     <TM 1, BN 128>  M <= 128.
//   * warp 0: TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B) through a 4-deep mbarrier ring of 32/48 KB stages; the WEIGHT
//     tiles of the first ring fill are issued BEFORE griddepcontrol.wait (weights never depend on the previous kernel);
//   * warp 1: ONE elected thread issues tcgen05.mma.cta_group::1.kind::f16 with shared-memory descriptors; accumulators live in
//     TMEM, double buffered (2 x TM x 128 columns): tcgen05.commit releases smem stages and hands a finished accumulator to the
//     epilogue through mbarriers, and the mainloop of tile i+1 runs while tile i is being written out;
//   * warps 4..7: epilogue — tcgen05.ld (32x32b: one TMEM lane = one output row per thread), bias / residual / SwiGLU (tile = 64
//     gate + 64 up columns of the fused gate_up weight), bf16 stores; or, for the row-parallel sequence-parallel prefill, the
//     fused REDUCE-SCATTER epilogue (see GemmParams::rs_*).
// reference kernels replaced: K3 qkv, K4 output_projection_cte, K5 mlp (prefill variants), K15 collective matmul.
#include <algorithm>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int GM_BM = 128, GM_BK = 64, GM_STAGES = 4;
constexpr int GM_THREADS = 256;
constexpr int GM_A_BYTES = GM_BM * GM_BK * 2;   // one 128-row A sub-tile

struct GemmParams {
  CUtensorMap tma_a;  // A [M,K]: dims {K, M}, box {64, TM*128}, SWIZZLE_128B (one instruction per k-block for both sub-tiles)
  CUtensorMap tma_b;  // B [N,K]: dims {K, N}, box {64, 128 (64 for GLU)}
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* residual;  // [M, N_out] or null
  __nv_bfloat16* c;               // [M, N_out]
  int M, N, K, ldc, act;          // act: 0 none, 1 silu*up, 2 gelu_tanh*up, 3 gelu*up, 4 GPT-OSS clamped SwiGLU  (N_out = N/2 when act != 0)
  int m_tiles, n_tiles;           // tile grid (CTA tiles of TM*128 x 128 output columns; 64 output features for GLU)
  // fp8 (e4m3 x e4m3 -> fp32, tcgen05 kind::f8f6f4): acc * a_scale[row] * w_scale[col] (dynamic per-token activation scale,
  // per-output-channel or per-tensor weight scale); null for bf16
  const float* a_scale;
  const float* w_scale;
  int w_scale_n;
  int splits;                     // split-K factor: work item = (tile, k slice); partial accumulators meet in `ws`
  float* ws;                      // [tiles][splits][TM*128][128] fp32 partials (L2 resident)
  unsigned* tickets;              // [tiles], self-resetting
  // ---- fused reduce-scatter epilogue (row-parallel layers of the sequence-parallel prefill; rs_world == 0: plain GEMM) ----
  // `c` is then this rank's symmetric staging buffer of PARTIAL sums.  After a tile's stores the CTA publishes the collective's
  // tag into slot [tile][my rank] of the rank that OWNS the tile's rows (st.release.sys over the peer mapping).  When a CTA has
  // finished its own tiles its epilogue warps turn to the tiles this rank owns: they wait until all `rs_world` slots of a tile
  // carry the tag, pull the tile REDUCED IN THE SWITCH (multimem.ld_reduce over the multicast mapping of the staging buffer),
  // add the residual and write the private output rows.  The transfer of a tile thus overlaps the GEMM of the following ones.
  uint32_t* rs_flags[SYMM_MAX_RANKS];   // [tiles][world] u32 on every rank (peer mapped)
  const uint8_t* rs_mc;                 // multicast address of the staging buffer
  __nv_bfloat16* rs_out;                // [owned rows][N] private output
  const __nv_bfloat16* rs_residual;     // [owned rows][N] or null
  const uint32_t* rs_step;              // device step counter (tag = (step << 8 | call) + 1)
  int rs_call, rs_rank, rs_world;
  int rs_tiles_per_seg;                 // m-tiles per segment (batch row); rank r owns tiles [r, r+1) * tiles_per_seg / world of each
  // ---- all-reduce flavour of the same epilogue: the reduced tile (+ residual, indexed by GLOBAL row) is multicast into every
  // rank's copy of a symmetric output buffer (multimem.st); the kernel ends with a cross-GPU completion flag exchange, so when it
  // has finished on a rank the whole [M, N] result is present there.  rs_bcast_mc == null: reduce-scatter into rs_out.
  uint8_t* rs_bcast_mc;                 // multicast address of the symmetric OUTPUT buffer [M, N]
  uint32_t* rs_done[SYMM_MAX_RANKS];    // [world] u32 on every rank (peer mapped): "rank s has broadcast all its tiles"
  unsigned* rs_cta_counter;             // local: CTAs of this launch that finished their reduce tiles (self-resetting)
  // ---- grouped (mixture-of-experts) flavour: A holds the tokens PERMUTED by expert, every expert's rows padded to whole 128-row
  // tiles (csrc/moe_grouped.cu builds the permutation on the device); B is the stacked weight [E * N, K] and m-tile `mt` multiplies
  // with the rows of expert grp_tile_expert[mt] (< 0: tile not in use this step, skipped by every role).  M is the STATIC upper
  // bound of permuted rows, so the launch is shape-stable under CUDA graphs whatever the routing.
  const int* grp_tile_expert;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n));
}
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(s_u32(b)), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  while (!mb_try(b, parity)) {
  }
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   s_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(s_u32(bar))
               : "memory");
}
// UMMA shared-memory descriptor: K-major tile, 128-byte rows, SWIZZLE_128B; 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(const void* smem) {
  const uint32_t addr = s_u32(smem);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows x 128 B             [32,46)
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)                  [46,48)
  d |= (uint64_t)2 << 61;                   // layout type SWIZZLE_128B                         [61,64)
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N at [17,23) (N>>3), M at [24,29) (M>>4)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor for kind::f8f6f4: D = f32, A = B = e4m3 (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f8(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

__device__ __forceinline__ void mb_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}

template <int TM, int BN, bool FP8>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  constexpr int B_BYTES = BN * GM_BK * 2;
  constexpr int STAGE_BYTES = TM * GM_A_BYTES + B_BYTES;
  constexpr int ACC_COLS = TM * BN;            // TMEM columns of one accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;         // double buffered: 256 (TM=1) or 512 (TM=2)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (s_u32(smem_raw) & 1023u)) & 1023u);
  // stage s: [TM][128][64] A sub-tiles, then [128][64] B tile (all 128B-swizzled, 1024-aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GM_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GM_STAGES;
  uint64_t* tmem_full = empty_bar + GM_STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool glu = p.act != 0;
  constexpr int BKE = FP8 ? 128 : 64;   // elements per 128-byte k-block row
  const int nkb = p.K / BKE;
  const int S = p.splits;
  const int total_tiles = p.m_tiles * p.n_tiles * S;   // work items: (tile, k slice), k slice fastest
  const int tile_out = glu ? BN / 2 : BN;   // output columns per tile

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GM_STAGES; ++s) {
      mb_init(&full_bar[s], 2);    // two producers (weights: warp 0, activations: warp 3), each posts its own byte count
      mb_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mb_init(&tmem_full[b], 1);
      mb_init(&tmem_empty[b], 4);   // the four epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();

  if (warp == 0 || warp == 3) {
    // ================= TMA producers =================
    // One issuing thread sustains ~one TMA instruction per 0.2 us (tools/bench_stream.cu); a k-block of 48 KB issued as three
    // instructions by ONE thread took 0.5 us while the tensor pipe needs 0.26 us for it.  So the two operands have their own
    // producer warp — warp 0 streams the WEIGHT tile (and starts before griddepcontrol.wait: weights do not depend on the previous
    // kernel), warp 3 the ACTIVATION tile (one 256-row box when TM = 2) — both post to the same full barrier.
    if (lane == 0) {
      const bool is_b = warp == 0;
      // (grouped: the routing plan comes from the preceding kernels, so the weight producer waits too)
      if (!is_b || p.grp_tile_expert != nullptr) pdl_wait();
      long long it = 0;   // k-block counter across work items (ring position)
      for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
        const int t = w / S, ks = w % S;
        const int m0 = (t % p.m_tiles) * (TM * GM_BM), n_out0 = (t / p.m_tiles) * tile_out;
        const int kb0 = (int)((long long)nkb * ks / S), kb1 = (int)((long long)nkb * (ks + 1) / S);
        int eo = 0;       // grouped: first row of this tile's expert inside the stacked weight
        if (p.grp_tile_expert != nullptr) {
          const int e = __ldg(p.grp_tile_expert + t % p.m_tiles);
          if (e < 0) continue;
          eo = e * p.N;
        }
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % GM_STAGES);
          const uint32_t ph = (uint32_t)((it / GM_STAGES) & 1);
          mb_wait(&empty_bar[s], ph ^ 1u);
          if (is_b) {
            mb_expect(&full_bar[s], B_BYTES);
            uint8_t* sB = smem + s * STAGE_BYTES + TM * GM_A_BYTES;
            if (glu) {  // 64 gate rows + 64 up rows of the fused [gate; up] weight
              tma_2d(sB, &p.tma_b, kb * BKE, eo + n_out0, &full_bar[s]);
              tma_2d(sB + B_BYTES / 2, &p.tma_b, kb * BKE, eo + (p.N >> 1) + n_out0, &full_bar[s]);
            } else {
              tma_2d(sB, &p.tma_b, kb * BKE, eo + n_out0, &full_bar[s]);
            }
          } else {
            mb_expect(&full_bar[s], TM * GM_A_BYTES);
            tma_2d(smem + s * STAGE_BYTES, &p.tma_a, kb * BKE, m0, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = FP8 ? umma_idesc_f8(GM_BM, BN) : umma_idesc(GM_BM, BN);
      long long it = 0;
      int tc = 0;
      if (p.grp_tile_expert != nullptr) pdl_wait();
      for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
        const int ks = w % S;
        if (p.grp_tile_expert != nullptr && __ldg(p.grp_tile_expert + (w / S) % p.m_tiles) < 0) continue;
        const int kb0 = (int)((long long)nkb * ks / S), kb1 = (int)((long long)nkb * (ks + 1) / S);
        const int buf = tc & 1;
        mb_wait(&tmem_empty[buf], (uint32_t)(((tc >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % GM_STAGES);
          const uint32_t ph = (uint32_t)((it / GM_STAGES) & 1);
          mb_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t db = umma_desc(smem + s * STAGE_BYTES + TM * GM_A_BYTES);
#pragma unroll
          for (int sub = 0; sub < TM; ++sub) {
            const uint64_t da = umma_desc(smem + s * STAGE_BYTES + sub * GM_A_BYTES);
            const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS + sub * BN);
#pragma unroll
            for (int k = 0; k < GM_BK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              // (fp8: 32 elements = 32 bytes per instruction; bf16: 16 elements = 32 bytes — the same descriptor step)
              if (FP8) umma_f8(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb != kb0 || k != 0) ? 1u : 0u);
              else umma_f16(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);       // frees the smem stage when these MMAs retire
        }
        umma_commit(&tmem_full[buf]);       // accumulator(s) of this tile complete
        ++tc;
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> (split-K fix-up) -> global =================
    pdl_wait();   // residual rows come from the previous kernels
    __shared__ int s_last;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int etid = threadIdx.x - 128;     // 0..127 inside the epilogue group
    const int n_out = glu ? (p.N >> 1) : p.N;
    // final math + store of 32 accumulator columns [c0, c0+32) of one row (GLU: gate chunk + matching up chunk)
    int cur_eo = 0;   // grouped: offset of the current tile's expert inside bias / w_scale ([E][N])
    auto store_chunk = [&](int row, int n_out0, int c0, const uint32_t (&vg)[32], const uint32_t (&vu)[32]) {
      const int col0 = n_out0 + c0;
      if (row >= p.M || col0 >= n_out) return;
      __nv_bfloat16* dst = p.c + (size_t)row * p.ldc + col0;
      const __nv_bfloat16* res = (!glu && p.residual) ? p.residual + (size_t)row * p.ldc + col0 : nullptr;
      // fp8: combined dequantisation factors of this row x these 32 columns (row scale x channel scale), vector loads
      float sc_g[32], sc_u[32];
      if (FP8) {
        const float as = p.a_scale[row];
        const bool full = col0 + 32 <= n_out && p.w_scale_n != 1;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 g4, u4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (full) {
            g4 = __ldg(reinterpret_cast<const float4*>(p.w_scale + cur_eo + col0 + j));
            if (glu) u4 = __ldg(reinterpret_cast<const float4*>(p.w_scale + cur_eo + n_out + col0 + j));
          } else if (p.w_scale_n == 1) {
            g4 = u4 = make_float4(p.w_scale[0], p.w_scale[0], p.w_scale[0], p.w_scale[0]);
          } else {
            float t[4], tu[4];
            for (int e = 0; e < 4; ++e) {
              const bool in = col0 + j + e < n_out;
              t[e] = in ? p.w_scale[cur_eo + col0 + j + e] : 0.f;
              tu[e] = (in && glu) ? p.w_scale[cur_eo + n_out + col0 + j + e] : 0.f;
            }
            g4 = make_float4(t[0], t[1], t[2], t[3]);
            u4 = make_float4(tu[0], tu[1], tu[2], tu[3]);
          }
          sc_g[j] = g4.x * as; sc_g[j + 1] = g4.y * as; sc_g[j + 2] = g4.z * as; sc_g[j + 3] = g4.w * as;
          sc_u[j] = u4.x * as; sc_u[j + 1] = u4.y * as; sc_u[j + 2] = u4.z * as; sc_u[j + 3] = u4.w * as;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = __uint_as_float(vg[j + e]);
          const bool in = col0 + j + e < n_out;
          float up_s = 1.f;
          if (FP8) {
            a *= sc_g[j + e];
            up_s = sc_u[j + e];
          }
          if (p.bias && in) a += __bfloat162float(p.bias[cur_eo + col0 + j + e]);
          if (glu) {
            float up = __uint_as_float(vu[j + e]) * up_s;
            if (p.bias && in) up += __bfloat162float(p.bias[cur_eo + n_out + col0 + j + e]);
            if (p.act == 4) {   // GPT-OSS clamped SwiGLU: (clamp(up, -7, 7) + 1) * g * sigmoid(1.702 g), g = min(gate, 7)
              const float g = fminf(a, 7.f), u = fminf(fmaxf(up, -7.f), 7.f);
              a = (u + 1.f) * g * (1.f / (1.f + __expf(-1.702f * g)));
            } else {
              a = (p.act == 1 ? silu(a) : (p.act == 2 ? gelu_tanh(a) : gelu_erf(a))) * up;
            }
          }
          f[e] = a;
        }
        if (col0 + j + 8 <= n_out) {
          if (res) {
            const uint4 r = ldg_act(res + j);
            f[0] += bf16lo(r.x); f[1] += bf16hi(r.x); f[2] += bf16lo(r.y); f[3] += bf16hi(r.y);
            f[4] += bf16lo(r.z); f[5] += bf16hi(r.z); f[6] += bf16lo(r.w); f[7] += bf16hi(r.w);
          }
          *reinterpret_cast<uint4*>(dst + j) =
              make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        } else {
          for (int e = 0; e < 8; ++e)
            if (col0 + j + e < n_out) dst[j + e] = __float2bfloat16(f[e] + (res ? ldg_act_bf16(res + j + e) : 0.f));
        }
      }
    };
    const int half = glu ? BN / 2 : BN;   // accumulator columns holding "gate or value"
    int tc = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
      const int t = w / S, ks = w % S;
      if (p.grp_tile_expert != nullptr) {
        const int e = __ldg(p.grp_tile_expert + t % p.m_tiles);
        if (e < 0) continue;
        cur_eo = e * p.N;
      }
      const int buf = tc & 1;
      const int m0 = (t % p.m_tiles) * (TM * GM_BM), n_out0 = (t / p.m_tiles) * tile_out;
      mb_wait(&tmem_full[buf], (uint32_t)((tc >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* my_ws = S > 1 ? p.ws + ((size_t)t * S + ks) * (TM * GM_BM) * BN : nullptr;
#pragma unroll 1
      for (int sub = 0; sub < TM; ++sub) {
        const int r_in = sub * GM_BM + q * 32 + lane;         // row inside the CTA tile (one per thread)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC_COLS + sub * BN);
        if (S > 1) {
          // split-K: park the fp32 partial tile (row-contiguous 128 B chunks per thread)
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint4* d4 = reinterpret_cast<uint4*>(my_ws + (size_t)r_in * BN + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) d4[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < half; c0 += 32) {
            uint32_t vg[32], vu[32];
            tmem_ld32(lane_addr + c0, vg);
            if (glu) tmem_ld32(lane_addr + BN / 2 + c0, vu);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            store_chunk(m0 + r_in, n_out0, c0, vg, vu);
          }
        }
      }
      // this warp has read its lanes of the buffer: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(&tmem_empty[buf]);
      if (p.rs_world > 0) {
        // publish: this rank's partial tile is in its staging buffer (barrier + system-scope release by one thread)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          const int mt = t % p.m_tiles;
          const int per = p.rs_tiles_per_seg / p.rs_world;
          const int owner = (mt % p.rs_tiles_per_seg) / per;
          uint32_t* fl = p.rs_flags[0];
#pragma unroll
          for (int d = 1; d < SYMM_MAX_RANKS; ++d)
            if (d == owner) fl = p.rs_flags[d];
          __threadfence_system();
          st_release_sys(fl + (size_t)t * p.rs_world + p.rs_rank, ll_tag(p.rs_step, p.rs_call));
        }
      }
      if (S > 1) {
        // the LAST k slice of a tile to arrive sums the partials (slice order: deterministic) and runs the real epilogue
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) s_last = (atomicAdd(&p.tickets[t], 1u) == (unsigned)(S - 1)) ? 1 : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (s_last) {
          __threadfence();
          const float* tile_ws = p.ws + (size_t)t * S * (TM * GM_BM) * BN;
#pragma unroll 1
          for (int sub = 0; sub < TM; ++sub) {
            const int r_in = sub * GM_BM + q * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < half; c0 += 32) {
              uint32_t vg[32], vu[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) vg[j] = vu[j] = 0u;
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {      // S <= 4: fully unrolled so that all partial loads of a chunk are in flight
                if (sl >= S) break;
                const float* src = tile_ws + ((size_t)sl * (TM * GM_BM) + r_in) * BN + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 a = __ldcg(reinterpret_cast<const float4*>(src + j));
                  vg[j] = __float_as_uint(__uint_as_float(vg[j]) + a.x);
                  vg[j + 1] = __float_as_uint(__uint_as_float(vg[j + 1]) + a.y);
                  vg[j + 2] = __float_as_uint(__uint_as_float(vg[j + 2]) + a.z);
                  vg[j + 3] = __float_as_uint(__uint_as_float(vg[j + 3]) + a.w);
                  if (glu) {
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(src + BN / 2 + j));
                    vu[j] = __float_as_uint(__uint_as_float(vu[j]) + b.x);
                    vu[j + 1] = __float_as_uint(__uint_as_float(vu[j + 1]) + b.y);
                    vu[j + 2] = __float_as_uint(__uint_as_float(vu[j + 2]) + b.z);
                    vu[j + 3] = __float_as_uint(__uint_as_float(vu[j + 3]) + b.w);
                  }
                }
              }
              store_chunk(m0 + r_in, n_out0, c0, vg, vu);
            }
          }
          if (etid == 0) p.tickets[t] = 0;   // re-armed for the next launch / graph replay
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_last reusable
      }
      ++tc;
    }
    if (p.rs_world > 0) {
      // ---- reduce phase: the tiles whose rows this rank owns, dealt round-robin to the CTAs ----
      const uint32_t tag = ll_tag(p.rs_step, p.rs_call);
      const int per = p.rs_tiles_per_seg / p.rs_world;           // owned m-tiles per segment
      const int segs = p.m_tiles / p.rs_tiles_per_seg;
      const int owned_m = per * segs;
      const uint32_t* my_flags = p.rs_flags[0];
#pragma unroll
      for (int d = 1; d < SYMM_MAX_RANKS; ++d)
        if (d == p.rs_rank) my_flags = p.rs_flags[d];
      for (int u = blockIdx.x; u < owned_m * p.n_tiles; u += gridDim.x) {
        const int lm = u % owned_m, nt = u / owned_m;              // local m-tile (row block of the private output), n-tile
        const int mt = (lm / per) * p.rs_tiles_per_seg + p.rs_rank * per + lm % per;   // global m-tile
        const int t = nt * p.m_tiles + mt;
        if (etid < p.rs_world) {
          const long long t0 = clock64();
          // tags grow monotonically, and a peer that is already one collective ahead has certainly finished this one
          while ((int)(ld_acquire_sys(my_flags + (size_t)t * p.rs_world + etid) - tag) < 0) {
            if (clock64() - t0 > 8000000000LL) {
              printf("gemm reduce-scatter: rank %d timed out waiting for rank %d tile %d (tag %u)\n", p.rs_rank, etid, t, tag);
              __trap();
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 128 rows x BN columns of bf16, 16 bytes per access: thread -> (row, 16-byte column chunk)
        constexpr int CPR = BN / 8;                               // 16-byte chunks per row
        const int n0 = nt * BN;
        // 8 in-switch reductions in flight per thread (one is a ~3 us round trip through the NVSwitch)
        for (int i0 = etid; i0 < GM_BM * CPR; i0 += 128 * 8) {
          uint32_t q[8][4];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * 128;
            const int r = i / CPR, col = n0 + (i % CPR) * 8;
            if (i < GM_BM * CPR && col < p.N) {
              const size_t src = ((size_t)(mt * GM_BM + r) * p.ldc + col) * 2;      // byte offset inside the staging buffer
              asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                           : "=r"(q[j][0]), "=r"(q[j][1]), "=r"(q[j][2]), "=r"(q[j][3]) : "l"(p.rs_mc + src) : "memory");
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * 128;
            const int r = i / CPR, col = n0 + (i % CPR) * 8;
            if (i < GM_BM * CPR && col < p.N) {
              const bool bcast = p.rs_bcast_mc != nullptr;
              // reduce-scatter: rows of the private output; all-reduce: rows of the full tensor
              const size_t dsto = (size_t)((bcast ? mt : lm) * GM_BM + r) * p.N + col;
              if (p.rs_residual != nullptr) {
                const uint4 rr = ldg_act(p.rs_residual + dsto);
                q[j][0] = pack_bf16(bf16lo(q[j][0]) + bf16lo(rr.x), bf16hi(q[j][0]) + bf16hi(rr.x));
                q[j][1] = pack_bf16(bf16lo(q[j][1]) + bf16lo(rr.y), bf16hi(q[j][1]) + bf16hi(rr.y));
                q[j][2] = pack_bf16(bf16lo(q[j][2]) + bf16lo(rr.z), bf16hi(q[j][2]) + bf16hi(rr.z));
                q[j][3] = pack_bf16(bf16lo(q[j][3]) + bf16lo(rr.w), bf16hi(q[j][3]) + bf16hi(rr.w));
              }
              if (bcast)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.rs_bcast_mc + dsto * 2), "r"(q[j][0]),
                             "r"(q[j][1]), "r"(q[j][2]), "r"(q[j][3]) : "memory");
              else
                *reinterpret_cast<uint4*>(p.rs_out + dsto) = make_uint4(q[j][0], q[j][1], q[j][2], q[j][3]);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (p.rs_bcast_mc != nullptr) {
        // completion: the last CTA of this rank to finish its tiles tells every rank; nobody leaves before all ranks have told
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          __threadfence_system();
          if (atomicAdd(p.rs_cta_counter, 1u) == gridDim.x - 1) {
            *p.rs_cta_counter = 0;
            __threadfence_system();
#pragma unroll
            for (int d = 0; d < SYMM_MAX_RANKS; ++d)
              if (d < p.rs_world) st_release_sys(p.rs_done[d] + p.rs_rank, tag);
          }
        }
        if (etid < p.rs_world) {
          const uint32_t* mine = p.rs_done[0];
#pragma unroll
          for (int d = 1; d < SYMM_MAX_RANKS; ++d)
            if (d == p.rs_rank) mine = p.rs_done[d];
          const long long t0 = clock64();
          while ((int)(ld_acquire_sys(mine + etid) - tag) < 0) {
            if (clock64() - t0 > 8000000000LL) {
              printf("gemm all-reduce: rank %d timed out waiting for the completion flag of rank %d (tag %u)\n", p.rs_rank, etid, tag);
              __trap();
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) != cudaSuccess || q == nullptr)
      throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn2>(q);
  }
  return fn;
}
static void make_2d(CUtensorMap* tm, const void* ptr, int rows, int K, int ld_elems, int box_rows, bool fp8 = false) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * (fp8 ? 1 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(fp8 ? 128 : GM_BK), (cuuint32_t)box_rows};   // 128-byte rows either way
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(tm, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (gemm) failed: " + std::to_string((int)r));
}

// split-K scratch (per device): fp32 partial tiles + tickets
static size_t gemm_ws_bytes() { return (size_t)96 << 20; }
static int gemm_max_tickets() { return 1 << 16; }
static float* g_gemm_ws[16] = {nullptr};
static unsigned* g_gemm_tk[16] = {nullptr};
static float* gemm_ws() {
  int dev;
  cudaGetDevice(&dev);
  if (g_gemm_ws[dev] == nullptr) {
    if (cudaMalloc(&g_gemm_ws[dev], gemm_ws_bytes()) != cudaSuccess) throw std::runtime_error("gemm: split-K workspace allocation failed");
  }
  return g_gemm_ws[dev];
}
static unsigned* gemm_tickets() {
  int dev;
  cudaGetDevice(&dev);
  if (g_gemm_tk[dev] == nullptr) {
    if (cudaMalloc(&g_gemm_tk[dev], gemm_max_tickets() * sizeof(unsigned)) != cudaSuccess) throw std::runtime_error("gemm: ticket allocation failed");
    cudaMemset(g_gemm_tk[dev], 0, gemm_max_tickets() * sizeof(unsigned));
  }
  return g_gemm_tk[dev];
}

static void gemm_launch_impl(const void* a, int lda, const void* b, const void* bias, const void* residual, void* c, int ldc, int M, int N,
                             int K, int act, cudaStream_t stream, const GemmRsArgs* rs, const float* a_scale, const float* w_scale,
                             int w_scale_n, const int* tile_expert = nullptr, int n_experts = 1) {
  GemmParams p{};
  p.grp_tile_expert = tile_expert;
  const bool glu = act != 0;
  const bool fp8 = a_scale != nullptr;
  p.a_scale = a_scale; p.w_scale = w_scale; p.w_scale_n = w_scale_n;
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.act = act;
  // TM = 2 (256-row CTA tiles) whenever there is more than one 128-row block: each weight tile then feeds two UMMAs
  static int force_tm = -1;
  if (force_tm < 0) {
    const char* e = getenv("NXDI_B200_GEMM_TM");
    force_tm = e ? atoi(e) : 0;
  }
  // tile shape: see the header.  Skinny problems (M <= 256) that would leave SMs without a tile trade weight reuse for occupancy:
  // first 128-row tiles (TM 1), then 64-column tiles — aggregate TMA throughput scales with the number of busy SMs
  // (M = 256, N = 6144: 48 tiles of 256x128 took 48 us, 192 tiles of 128x64 ~30 us; cuBLAS 23.5 us)
  int TM = (M > GM_BM && M <= 2 * GM_BM) ? 2 : 1;
  int BN = (TM == 1 && M > 2 * GM_BM) ? 256 : 128;
  int model_split = 0;
  if (M <= 4 * GM_BM) {
    // Weight-dominated sizes (M <= 512): a CTA is bound by its shared-memory fill rate (~100 GB/s per SM), so a tile costs
    // (TM*128 + BN) operand rows per k-block and the launch costs  waves x that  — pick the shape that minimises it (ties: the larger
    // tile).  Measured (profiles/r2/skinny_gemm_sweep.txt): M = 256 qkv 51 -> 37 us with 128x128 instead of 128x64 tiles (192 tiles were
    // two waves), o_proj best with 128x64 (one wave of 128 tiles).  Split-K only pays for very deep K with less than half the SMs
    // busy (down: 84 -> 68 us with two k slices); everywhere else the partial round trip costs more than it gains.
    const int n_out_ = glu ? N / 2 : N;
    const int cand[4][2] = {{2, 128}, {1, 256}, {1, 128}, {1, 64}};
    long best_cost = -1;
    for (int ci = 0; ci < 4; ++ci) {
      const int tm = cand[ci][0], bn = cand[ci][1];
      const int t_out = glu ? bn / 2 : bn;
      const long tiles_ = (long)((M + tm * GM_BM - 1) / (tm * GM_BM)) * ((n_out_ + t_out - 1) / t_out);
      const long cost = ((tiles_ + n_sms - 1) / n_sms) * (tm * GM_BM + bn);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; TM = tm; BN = bn; }
    }
    if (K >= 8192) {   // deep K: 128x128 tiles in two k slices when that still fits one wave
      const long t128 = (long)((M + GM_BM - 1) / GM_BM) * ((n_out_ + (glu ? 64 : 128) - 1) / (glu ? 64 : 128));
      if (2 * t128 <= n_sms && (GM_BM + 128) / 2 * 14 / 10 < best_cost) { TM = 1; BN = 128; model_split = 2; }
    }
  }
  if (force_tm) TM = force_tm;
  {
    // tuning override: NXDI_B200_GEMM_BN = 64 | 128 | 256 (256 only with 128-row tiles; 64 only with 128-row tiles)
    const char* e = getenv("NXDI_B200_GEMM_BN");
    const int fbn = e ? atoi(e) : 0;
    if (fbn == 64 || fbn == 128 || fbn == 256) {
      BN = fbn;
      if (BN != 128) TM = 1;
    }
  }
  if (tile_expert != nullptr) {   // grouped: one expert per 128-row tile
    if (M % GM_BM != 0) throw std::runtime_error("grouped gemm: the permuted row bound must be a multiple of 128");
    TM = 1;
    BN = 256;
  }
  if (rs != nullptr) {   // fused reduce-scatter: 128-row tiles so that a tile has exactly one owner; no split-K
    if (glu) throw std::runtime_error("gemm: the fused reduce-scatter needs a plain epilogue");
    TM = 1;
    BN = M > 2 * GM_BM ? 256 : 128;
  }
  make_2d(&p.tma_a, a, M, K, lda, TM * GM_BM, fp8);
  make_2d(&p.tma_b, b, N * n_experts, K, K, glu ? BN / 2 : BN, fp8);
  const int n_out = glu ? N / 2 : N;
  const int tile_out = glu ? BN / 2 : BN;
  p.m_tiles = (M + TM * GM_BM - 1) / (TM * GM_BM);
  p.n_tiles = (n_out + tile_out - 1) / tile_out;
  // split-K when the tile grid leaves SMs idle or quantises badly (skinny prefill: M <= 256, N of a few thousand):
  // minimise  ceil(tiles * S / SMs) / S  (+ a small fix-up charge), bounded by the workspace and >= 8 k-blocks per slice
  const int tiles = p.m_tiles * p.n_tiles, nkb = K / (fp8 ? 128 : GM_BK);
  static int force_s = -1;
  if (force_s < 0) {
    const char* e = getenv("NXDI_B200_GEMM_SPLITK");
    force_s = e ? atoi(e) : 0;
  }
  int best_s = 1;
  double best = 1e30;
  const size_t tile_ws_bytes = (size_t)TM * GM_BM * BN * 4;
  // only when at most a quarter of the SMs would get a tile (measured: the partial round trip costs ~10 us, so a 1.5-wave grid
  // such as gate_up at M = 256 — 224 tiles — is better left alone: 86 us unsplit vs 280 us with 3 slices)
  if (tiles * 4 <= n_sms) {
    for (int sgl = 1; sgl <= 4; ++sgl) {
      if (sgl > 1 && (nkb / sgl < 8 || (size_t)tiles * sgl * tile_ws_bytes > gemm_ws_bytes())) break;
      const double cost = (double)((tiles * sgl + n_sms - 1) / n_sms) / sgl + (sgl > 1 ? 0.06 : 0.0);
      if (cost < best - 1e-9) { best = cost; best_s = sgl; }
    }
  }
  if (model_split > 0 && force_s <= 0 && !force_tm) best_s = model_split;
  p.splits = force_s > 0 ? std::min({force_s, 4, std::max(1, nkb / 2)}) : best_s;
  if (force_s > 0 && (size_t)tiles * p.splits * tile_ws_bytes > gemm_ws_bytes()) p.splits = 1;
  if (rs != nullptr || tile_expert != nullptr) p.splits = 1;
  if ((size_t)tiles * p.splits * tile_ws_bytes > gemm_ws_bytes() || tiles > gemm_max_tickets()) p.splits = 1;
  p.ws = p.splits > 1 ? gemm_ws() : nullptr;
  p.tickets = gemm_tickets();
  if (rs != nullptr) {
    const int rows_per_seg = rs->rows_per_seg;
    if (rows_per_seg % (GM_BM * rs->world) != 0 || M % rows_per_seg != 0 || N % 8 != 0 || ldc != N)
      throw std::runtime_error("gemm: fused reduce-scatter needs rows_per_rank % 128 == 0 and a dense [M, N] staging buffer");
    for (int i = 0; i < rs->world; ++i) p.rs_flags[i] = reinterpret_cast<uint32_t*>(rs->flag_ptrs[i]);
    p.rs_mc = reinterpret_cast<const uint8_t*>(rs->mc);
    p.rs_out = reinterpret_cast<__nv_bfloat16*>(rs->out);
    p.rs_residual = reinterpret_cast<const __nv_bfloat16*>(rs->residual);
    p.rs_step = reinterpret_cast<const uint32_t*>(rs->step);
    p.rs_call = rs->call; p.rs_rank = rs->rank; p.rs_world = rs->world;
    p.rs_tiles_per_seg = rows_per_seg / GM_BM;
    p.rs_bcast_mc = reinterpret_cast<uint8_t*>(const_cast<void*>(rs->bcast_mc));
    for (int i = 0; i < rs->world; ++i) p.rs_done[i] = reinterpret_cast<uint32_t*>(rs->done_ptrs[i]);
    p.rs_cta_counter = reinterpret_cast<unsigned*>(rs->cta_counter);
    if ((long long)p.m_tiles * p.n_tiles > rs->max_tiles) throw std::runtime_error("gemm: too many tiles for the reduce-scatter flag array");
  }
  const size_t smem = (size_t)GM_STAGES * (TM * GM_A_BYTES + BN * GM_BK * 2) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    const int max_smem = GM_STAGES * (2 * GM_A_BYTES + 128 * GM_BK * 2) + 256 + 1024;   // == 1 x A + 256-row B
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    configured = true;
  }
  const int grid = std::min(n_sms, p.m_tiles * p.n_tiles * p.splits);
  const dim3 g(grid), b3(GM_THREADS);
  if (fp8) {
    if (TM == 2) launch_pdl(gemm_tcgen05_kernel<2, 128, true>, g, b3, smem, stream, p);
    else if (BN == 256) launch_pdl(gemm_tcgen05_kernel<1, 256, true>, g, b3, smem, stream, p);
    else if (BN == 64) launch_pdl(gemm_tcgen05_kernel<1, 64, true>, g, b3, smem, stream, p);
    else launch_pdl(gemm_tcgen05_kernel<1, 128, true>, g, b3, smem, stream, p);
  } else {
    if (TM == 2) launch_pdl(gemm_tcgen05_kernel<2, 128, false>, g, b3, smem, stream, p);
    else if (BN == 256) launch_pdl(gemm_tcgen05_kernel<1, 256, false>, g, b3, smem, stream, p);
    else if (BN == 64) launch_pdl(gemm_tcgen05_kernel<1, 64, false>, g, b3, smem, stream, p);
    else launch_pdl(gemm_tcgen05_kernel<1, 128, false>, g, b3, smem, stream, p);
  }
}

void gemm_tcgen05_launch(const void* a, int lda, const void* b, const void* bias, const void* residual, void* c, int ldc, int M,
                         int N, int K, int act, cudaStream_t stream, const GemmRsArgs* rs) {
  gemm_launch_impl(a, lda, b, bias, residual, c, ldc, M, N, K, act, stream, rs, nullptr, nullptr, 0);
}

// grouped GEMM of the mixture-of-experts layers: a [R, K] tokens permuted by expert (R % 128 == 0), b [E, N, K] stacked expert
// weights, tile_expert [R / 128] (device; -1 = unused tile), bias [E, N] or null; fp8 when a_scale != null (w_scale [E, N])
void gemm_grouped_launch(const void* a, const void* b, const void* bias, void* c, int R, int N, int K, int n_experts, int act,
                         const int* tile_expert, const float* a_scale, const float* w_scale, cudaStream_t stream) {
  if (K % (a_scale ? 128 : 64) != 0) throw std::runtime_error("grouped gemm: K must be a multiple of 64 (128 for fp8)");
  const int n_out = act ? N / 2 : N;
  gemm_launch_impl(a, K, b, bias, nullptr, c, n_out, R, N, K, act, stream, nullptr, a_scale, w_scale, a_scale ? N : 0, tile_expert,
                   n_experts);
}

// W8A8 fp8-e4m3 GEMM: a [M, K] fp8 with per-row scale, b [N, K] fp8 with per-channel (w_scale_n == N) or per-tensor (1) scale
void gemm_fp8_launch(const void* a, int lda, const void* b, const float* a_scale, const float* w_scale, int w_scale_n, const void* bias,
                     const void* residual, void* c, int ldc, int M, int N, int K, int act, cudaStream_t stream) {
  if (K % 128 != 0) throw std::runtime_error("gemm_fp8: K must be a multiple of 128");
  gemm_launch_impl(a, lda, b, bias, residual, c, ldc, M, N, K, act, stream, nullptr, a_scale, w_scale, w_scale_n);
}

}  // namespace nxdi
