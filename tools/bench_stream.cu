// Streaming micro-benchmark (design input for the decode GEMV): how fast can ONE SM / the whole chip pull weights into shared
// memory with (a) 3-D tiled TMA boxes of 128-byte rows, (b) contiguous 1-D bulk copies (pre-tiled weights), (c) LDG.128?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/bench_stream tools/bench_stream.cu -lcuda
//   ./bench_stream            -> table: variant, bytes per copy, ring KB, CTAs/SM, TB/s
// No math: the consumer releases a stage as soon as it is full, so the number is the data-movement ceiling.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }

struct P {
  CUtensorMap tm;
  const uint8_t* src;
  long long bytes_per_cta;
  int mode;        // 0: 3-D TMA box {64, kg, 16}; 1: 1-D bulk copy; 2: LDG.128
  int copy_bytes;  // bytes per copy instruction (TMA box / bulk size)
  int n_stages;
  int K;           // mode 0: row length (elements)
  int rows_per_cta;
  unsigned long long* sink;
};

// 1 producer warp (lane 0 issues) + 1 consumer warp (releases immediately).  mode 2: 8 warps of LDG.
__global__ void __launch_bounds__(320, 2) stream_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[64], empty[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = p.n_stages;
  if (p.mode == 2) {
    // LDG path: 8 warps x 32 lanes x 16 B, 8 loads in flight per lane, XOR-accumulate so the loads are not dead code
    if (warp >= 8) return;
    const uint4* src = reinterpret_cast<const uint4*>(p.src + (long long)blockIdx.x * p.bytes_per_cta);
    const long long n = p.bytes_per_cta / 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (long long i = tid; i + 7 * 256 < n; i += 8 * 256) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src + i + j * 256));
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) p.sink[0] = 1;
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long count = p.bytes_per_cta / p.copy_bytes;
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      const int kg = p.copy_bytes / (16 * 128);               // mode 0: 128-byte groups per row per box
      const int chunks = p.K / (64 * kg);                     // boxes per 16-row tile
      for (long long i = 0; i < count; ++i) {
        mbar_wait(&empty[stage], ph ^ 1u);
        mbar_expect(&full[stage], p.copy_bytes);
        uint8_t* dst = base + (size_t)stage * p.copy_bytes;
        if (p.mode == 0) {
          const int tile = (int)(i / chunks), chunk = (int)(i % chunks);
          const int row0 = blockIdx.x * p.rows_per_cta + tile * 16;
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(su32(dst)), "l"(&p.tm), "r"(0), "r"(chunk * kg), "r"(row0), "r"(su32(&full[stage])) : "memory");
        } else {
          const uint8_t* s = p.src + (long long)blockIdx.x * p.bytes_per_cta + i * p.copy_bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(su32(dst)), "l"(s), "r"(p.copy_bytes), "r"(su32(&full[stage])) : "memory");
        }
        if (++stage == NS) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t ph = 0;
    for (long long i = 0; i < count; ++i) {
      mbar_wait(&full[stage], ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NS) { stage = 0; ph ^= 1u; }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const long long total = 2LL << 30;   // 2 GiB >> L2
  uint8_t* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  const int K = 4096;
  const long long N = total / (K * 2);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("%-10s %8s %8s %6s %8s %10s\n", "variant", "copy B", "ring KB", "cta/SM", "us", "TB/s");
  struct Cfg { int mode, copy, ring_kb, per_sm; };
  Cfg cfgs[] = {
      {0, 4096, 84, 1}, {0, 8192, 84, 1}, {0, 16384, 96, 1}, {0, 4096, 192, 1}, {0, 8192, 192, 1}, {0, 16384, 192, 1},
      {0, 4096, 84, 2}, {0, 8192, 84, 2}, {0, 16384, 96, 2},
      {1, 2048, 84, 1}, {1, 4096, 84, 1}, {1, 8192, 84, 1}, {1, 16384, 96, 1}, {1, 32768, 96, 1},
      {1, 4096, 192, 1}, {1, 8192, 192, 1}, {1, 16384, 192, 1}, {1, 32768, 192, 1},
      {1, 4096, 84, 2}, {1, 8192, 84, 2}, {1, 16384, 96, 2},
      {2, 16, 0, 1}, {2, 16, 0, 2},
  };
  for (const Cfg& c : cfgs) {
    P p{};
    p.src = buf; p.mode = c.mode; p.copy_bytes = c.copy; p.K = K; p.sink = sink;
    const int grid = sms * c.per_sm;
    long long rows_per_cta = (N / grid) / 16 * 16;
    p.rows_per_cta = (int)rows_per_cta;
    p.bytes_per_cta = rows_per_cta * K * 2;
    p.n_stages = c.mode == 2 ? 1 : (c.ring_kb * 1024) / c.copy;
    if (p.n_stages > 64) p.n_stages = 64;
    if (c.mode == 0) {
      const int kg = c.copy / (16 * 128);
      cuuint64_t gdim[3] = {64, (cuuint64_t)(K / 64), (cuuint64_t)N};
      cuuint64_t gstr[2] = {128, (cuuint64_t)K * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)kg, 16};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    }
    const size_t smem = c.mode == 2 ? 1024 : (size_t)p.n_stages * c.copy + 2048;
    const int threads = c.mode == 2 ? 256 : 64;
    for (int it = 0; it < 2; ++it) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<grid, threads, smem>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tb = (double)p.bytes_per_cta * grid / (ms * 1e-3) / 1e12;
    printf("%-10s %8d %8d %6d %8.1f %10.3f\n", c.mode == 0 ? "tma3d" : (c.mode == 1 ? "bulk1d" : "ldg128"), c.copy,
           c.mode == 2 ? 0 : (int)((size_t)p.n_stages * c.copy / 1024), c.per_sm, ms * 1e3, tb);
    fflush(stdout);
  }
  // single-SM ceiling: one CTA streams 64 MB
  for (int mode = 0; mode < 2; ++mode) {
    for (int copy : {4096, 16384}) {
      P p{};
      p.src = buf; p.mode = mode; p.copy_bytes = copy; p.K = K; p.sink = sink;
      p.rows_per_cta = 8192; p.bytes_per_cta = 8192LL * K * 2;
      p.n_stages = (192 * 1024) / copy; if (p.n_stages > 64) p.n_stages = 64;
      if (mode == 0) {
        const int kg = copy / (16 * 128);
        cuuint64_t gdim[3] = {64, (cuuint64_t)(K / 64), (cuuint64_t)N};
        cuuint64_t gstr[2] = {128, (cuuint64_t)K * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)kg, 16};
        cuuint32_t es[3] = {1, 1, 1};
        enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      const size_t smem = (size_t)p.n_stages * copy + 2048;
      for (int it = 0; it < 2; ++it) {
        CK(cudaEventRecord(e0));
        stream_kernel<<<1, 64, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("ONE SM %-8s copy %6d ring %3d KB: %8.1f us  %7.1f GB/s\n", mode == 0 ? "tma3d" : "bulk1d", copy,
             (int)((size_t)p.n_stages * copy / 1024), ms * 1e3, (double)p.bytes_per_cta / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
