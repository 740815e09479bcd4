// Launchers for the weight-streaming skinny GEMM (see gemv.cuh) and the stand-alone RMSNorm.
#include <algorithm>

#include "gemv.cuh"

namespace nxdi {

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

size_t gemv_smem_bytes(int T, int K, bool x_in_smem) {
  return (x_in_smem ? (size_t)T * (K * 2 + 64) : 0) + (2 * GEMV_WARPS * 128 + 64) * sizeof(float);
}

template <bool GLU>
static void launch_gemv(const GemvParams& p, cudaStream_t stream) {
  auto kern = gemv_kernel<GLU>;
  const size_t smem = gemv_smem_bytes(p.T, p.K, p.x_in_smem);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured = true;
  }
  const int n_tiles = GLU ? ((p.N / 2) + 7) / 8 : (p.N + 15) / 16;
  const int per_sm = smem <= 112 * 1024 ? 2 : 1;
  const int grid = std::min(n_tiles, per_sm * num_sms());
  launch_pdl(kern, dim3(grid), dim3(GEMV_THREADS), smem, stream, p);
}

void gemv_launch(const GemvParams& p, int mode, cudaStream_t stream) {
  const bool glu = p.act != ACT_NONE;
  if (mode != 0) throw std::runtime_error("gemv v1: the fused all-reduce lives in gemv2 (activations must fit shared memory)");
  if (glu) launch_gemv<true>(p, stream);
  else launch_gemv<false>(p, stream);
}

// rmsnorm kernel (also used to pre-normalise x when T*K does not fit the GEMV's shared memory)
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ res_in,
                                                      const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y,
                                                      __nv_bfloat16* __restrict__ res_out, int H, float eps, float offset) {
  __shared__ float sred[8];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * H);
  const uint4* rr = res_in ? reinterpret_cast<const uint4*>(res_in + (size_t)row * H) : nullptr;
  uint4* ro = res_out ? reinterpret_cast<uint4*>(res_out + (size_t)row * H) : nullptr;
  const int nvec = H >> 3;
  constexpr int MAXV = 8;  // up to H = 16384 in registers
  float v[MAXV][8];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = tid + i * 256;
    if (idx < nvec) {
      uint4 q = xr[idx];
      float f[8] = {bf16lo(q.x), bf16hi(q.x), bf16lo(q.y), bf16hi(q.y), bf16lo(q.z), bf16hi(q.z), bf16lo(q.w), bf16hi(q.w)};
      if (rr) {
        uint4 r = rr[idx];
        float g[8] = {bf16lo(r.x), bf16hi(r.x), bf16lo(r.y), bf16hi(r.y), bf16lo(r.z), bf16hi(r.z), bf16lo(r.w), bf16hi(r.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j] + g[j]));  // residual stream is bf16
        if (ro) ro[idx] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = f[j];
        acc += f[j] * f[j];
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) sred[warp] = acc;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sred[i];
  const float rstd = rsqrtf(tot / (float)H + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * H);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = tid + i * 256;
    if (idx < nvec) {
      uint4 g = wr[idx];
      float gm[8] = {bf16lo(g.x), bf16hi(g.x), bf16lo(g.y), bf16hi(g.y), bf16lo(g.z), bf16hi(g.z), bf16lo(g.w), bf16hi(g.w)};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rstd * (gm[j] + offset);
      yr[idx] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

void rmsnorm_launch(const void* x, const void* res_in, const void* w, void* y, void* res_out, int rows, int H, float eps,
                    float offset, cudaStream_t stream) {
  launch_pdl(rmsnorm_kernel, dim3(rows), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
             reinterpret_cast<const __nv_bfloat16*>(res_in), reinterpret_cast<const __nv_bfloat16*>(w),
             reinterpret_cast<__nv_bfloat16*>(y), reinterpret_cast<__nv_bfloat16*>(res_out), H, eps, offset);
}

}  // namespace nxdi
