// Prefill / large-T GEMM on the 5th-generation tensor cores:  C[M,N] = epilogue( A[M,K] · B[N,K]^T )   (bf16 in, fp32 acc)
//
//   A = activations [tokens, K] (K-major), B = nn.Linear weight [out, K] (K-major)  ->  no transposes anywhere.
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128x64 A and 128x64 B tiles through a 4-deep mbarrier ring;
//   * ONE elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128x128x16) with shared-memory descriptors;
//     the 128x128 fp32 accumulator lives in TMEM (128 columns); tcgen05.commit releases smem stages / signals the
//     epilogue through mbarriers — no thread ever holds accumulator fragments during the main loop;
//   * 4 epilogue warps read the accumulator with tcgen05.ld (32x32b: one TMEM lane = one output row per thread),
//     apply bias / residual / SwiGLU (tile = 64 gate + 64 up columns of the fused gate_up weight) and store bf16.
//   Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (lane quarter = warp % 4).
// reference kernels replaced: K3 qkv, K4 output_projection_cte, K5 mlp (prefill variants), K15 collective matmul.
#include <algorithm>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

// GM_BN is a template parameter: 128 for large grids; 64 when a 128-wide tiling would leave SMs idle (prefill with a few
// hundred tokens: M/128 x N/128 tiles < 2 x #SMs) — twice the CTAs, 96 KB of shared memory so two CTAs share an SM.
constexpr int GM_BM = 128, GM_BK = 64, GM_STAGES = 4;
constexpr int GM_THREADS = 256;
constexpr int GM_A_BYTES = GM_BM * GM_BK * 2;

struct GemmParams {
  CUtensorMap tma_a;  // A [M,K]: dims {K, M}, box {64, 128}, SWIZZLE_128B
  CUtensorMap tma_b;  // B [N,K]: dims {K, N}, box {64, 128 (64 for GLU)}
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* residual;  // [M, N_out] or null
  __nv_bfloat16* c;               // [M, N_out]
  int M, N, K, ldc, act;          // act: 0 none, 1 silu*up, 2 gelu_tanh*up, 3 gelu*up  (N_out = N/2 when act != 0)
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

template <int GM_BN>
__global__ void __launch_bounds__(GM_THREADS, GM_BN == 64 ? 2 : 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  constexpr int GM_B_BYTES = GM_BN * GM_BK * 2;
  constexpr int GM_TMEM_COLS = GM_BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                  // [STAGES][128][64] bf16, swizzled
  uint8_t* sB = smem + GM_STAGES * GM_A_BYTES;         // [STAGES][128][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + GM_STAGES * GM_B_BYTES);
  uint64_t* empty_bar = full_bar + GM_STAGES;
  uint64_t* tmem_full = empty_bar + GM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // raster: M tiles vary fastest, so the CTAs resident at any moment share a handful of weight (B) tiles and stream the
  // weights from HBM exactly once; A (activations, a few MB) is what gets re-read, from L2.  (ncu on the first version,
  // N-fastest: 3.77 GB of DRAM reads for a 0.25 GB problem.)
  const int m0 = blockIdx.x * GM_BM;
  const bool glu = p.act != 0;
  const int n_out0 = blockIdx.y * (glu ? GM_BN / 2 : GM_BN);  // first output column of this tile
  const int nkb = p.K / GM_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GM_STAGES; ++s) {
      mb_init(&full_bar[s], 1);
      mb_init(&empty_bar[s], 1);
    }
    mb_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(GM_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  pdl_wait();  // A (activations) is produced by the previous kernel

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GM_STAGES;
        const uint32_t ph = (uint32_t)((kb / GM_STAGES) & 1);
        mb_wait(&empty_bar[s], ph ^ 1u);
        mb_expect(&full_bar[s], GM_A_BYTES + GM_B_BYTES);
        tma_2d(sA + s * GM_A_BYTES, &p.tma_a, kb * GM_BK, m0, &full_bar[s]);
        if (glu) {  // 64 gate rows + 64 up rows of the fused [gate; up] weight
          tma_2d(sB + s * GM_B_BYTES, &p.tma_b, kb * GM_BK, n_out0, &full_bar[s]);
          tma_2d(sB + s * GM_B_BYTES + GM_B_BYTES / 2, &p.tma_b, kb * GM_BK, (p.N >> 1) + n_out0, &full_bar[s]);
        } else {
          tma_2d(sB + s * GM_B_BYTES, &p.tma_b, kb * GM_BK, n_out0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(GM_BM, GM_BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GM_STAGES;
        const uint32_t ph = (uint32_t)((kb / GM_STAGES) & 1);
        mb_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = umma_desc(sA + s * GM_A_BYTES), db = umma_desc(sB + s * GM_B_BYTES);
#pragma unroll
        for (int k = 0; k < GM_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full);        // accumulator complete
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;     // one output row per thread
    mb_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_out = glu ? (p.N >> 1) : p.N;
    if (!glu) {
#pragma unroll 1
      for (int c0 = 0; c0 < GM_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int col0 = n_out0 + c0;
        if (row < p.M && col0 < n_out) {
          __nv_bfloat16* dst = p.c + (size_t)row * p.ldc + col0;
          const __nv_bfloat16* res = p.residual ? p.residual + (size_t)row * p.ldc + col0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = __uint_as_float(v[j + e]);
              if (p.bias) f[e] += __bfloat162float(p.bias[col0 + j + e]);
            }
            if (res) {
              const uint4 r = *reinterpret_cast<const uint4*>(res + j);
              f[0] += bf16lo(r.x); f[1] += bf16hi(r.x); f[2] += bf16lo(r.y); f[3] += bf16hi(r.y);
              f[4] += bf16lo(r.z); f[5] += bf16hi(r.z); f[6] += bf16lo(r.w); f[7] += bf16hi(r.w);
            }
            if (col0 + j + 8 <= n_out) {
              *reinterpret_cast<uint4*>(dst + j) =
                  make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            } else {
              for (int e = 0; e < 8; ++e)
                if (col0 + j + e < n_out) dst[j + e] = __float2bfloat16(f[e]);
            }
          }
        }
      }
    } else {
      // columns [0,64) = gate, [64,128) = up of the same 64 output features
#pragma unroll 1
      for (int c0 = 0; c0 < GM_BN / 2; c0 += 32) {
        uint32_t vg[32], vu[32];
        tmem_ld32(lane_addr + c0, vg);
        tmem_ld32(lane_addr + GM_BN / 2 + c0, vu);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int col0 = n_out0 + c0;
        if (row < p.M && col0 < n_out) {
          __nv_bfloat16* dst = p.c + (size_t)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float gte = __uint_as_float(vg[j + e]), up = __uint_as_float(vu[j + e]);
              if (p.bias) {
                gte += __bfloat162float(p.bias[col0 + j + e]);
                up += __bfloat162float(p.bias[n_out + col0 + j + e]);
              }
              const float a = p.act == 1 ? silu(gte) : (p.act == 2 ? gelu_tanh(gte) : gelu_erf(gte));
              f[e] = a * up;
            }
            if (col0 + j + 8 <= n_out) {
              *reinterpret_cast<uint4*>(dst + j) =
                  make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            } else {
              for (int e = 0; e < 8; ++e)
                if (col0 + j + e < n_out) dst[j + e] = __float2bfloat16(f[e]);
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(GM_TMEM_COLS));
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
static void make_2d(CUtensorMap* tm, const void* ptr, int rows, int K, int ld_elems, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {GM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (gemm) failed: " + std::to_string((int)r));
}

void gemm_tcgen05_launch(const void* a, int lda, const void* b, const void* bias, const void* residual, void* c, int ldc, int M,
                         int N, int K, int act, cudaStream_t stream) {
  GemmParams p;
  const bool glu = act != 0;
  make_2d(&p.tma_a, a, M, K, lda, GM_BM);
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_out_ = glu ? N / 2 : N;
  const long long tiles128 = (long long)((M + GM_BM - 1) / GM_BM) * ((n_out_ + (glu ? 63 : 127)) / (glu ? 64 : 128));
  static int force_bn = -1;
  if (force_bn < 0) {
    const char* e = getenv("NXDI_B200_GEMM_BN");
    force_bn = e ? atoi(e) : 0;
  }
  const int BN = force_bn ? force_bn : (tiles128 < 2LL * n_sms ? 64 : 128);
  make_2d(&p.tma_b, b, N, K, K, glu ? BN / 2 : BN);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.act = act;
  const size_t smem = GM_STAGES * (GM_A_BYTES + BN * GM_BK * 2) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(gemm_tcgen05_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         GM_STAGES * (GM_A_BYTES + 128 * GM_BK * 2) + 256 + 1024);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         GM_STAGES * (GM_A_BYTES + 64 * GM_BK * 2) + 256 + 1024);
    configured = true;
  }
  const int n_out = glu ? N / 2 : N;
  const int tile_n = glu ? BN / 2 : BN;
  dim3 grid((M + GM_BM - 1) / GM_BM, (n_out + tile_n - 1) / tile_n);
  if (BN == 64) launch_pdl(gemm_tcgen05_kernel<64>, grid, dim3(GM_THREADS), smem, stream, p);
  else launch_pdl(gemm_tcgen05_kernel<128>, grid, dim3(GM_THREADS), smem, stream, p);
}

}  // namespace nxdi
