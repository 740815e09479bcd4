// In-switch collectives over the symmetric heap (csrc/symm_heap.cpp): NVLS `multimem.*` kernels for the prefill-sized tensors
// of the tensor-parallel layers (the decode-sized ones ride the LL path fused into gemv2).
//
//   all-reduce      two-shot: rank r reduces slice r of the buffer IN THE SWITCH (multimem.ld_reduce, fp32 accumulate) and
//                   broadcasts the result to every copy (multimem.st); (+ residual, written to a private output)
//   reduce-scatter  rank r pulls the switch-reduced slice r (sequence-parallel row-parallel layers), + residual
//   all-gather      rank r multicasts its slice into every rank's buffer (sequence-parallel column-parallel layers)
// Every kernel brackets its data phase with a cross-GPU barrier on signal words of the heap: rank r stores the collective's tag
// into slot [cta][r] of every peer (st.release.sys over the peer mapping) and spins on its own slots (ld.acquire.sys, bounded).
// Reference call sites these replace: reduce_scatter / all_gather of models/model_base.py:1471-1583, all_reduce of the
// row-parallel layers, mappings of the external parallel_layers package.
#include <stdexcept>
#include <string>

#include "api.h"
#include "common.cuh"

namespace nxdi {

constexpr int NVLS_THREADS = 512;
constexpr int NVLS_MAX_CTAS = 64;

struct NvlsArgs {
  uint32_t* sig[SYMM_MAX_RANKS];   // peer-mapped signal words: [2 phase][NVLS_MAX_CTAS][world] u32 at the same heap offset
  const uint32_t* step;            // device step counter (tag = (step << 8 | call) + 1)
  int rank, world, call;
  // data
  uint8_t* mc;          // multicast address of the symmetric buffer
  uint8_t* local;       // this rank's address of the same buffer
  const __nv_bfloat16* residual;   // or null
  __nv_bfloat16* out;              // private output (all-reduce with residual / reduce-scatter) or null (in place)
  // layout: [segs][rows_per_seg][row_elems] bf16; rank r owns rows [r * rows_per_seg / world, (r+1) * rows_per_seg / world) of
  // EVERY segment (segment = batch row of a [B, T, H] activation sharded along T; a flat tensor is one segment)
  int segs, rows_per_seg, row_elems;
};

__device__ __forceinline__ void mm_ld_reduce_bf16x8(const void* mc_addr, uint32_t (&r)[4]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "l"(mc_addr)
               : "memory");
}
__device__ __forceinline__ void mm_st_16B(void* mc_addr, const uint32_t (&r)[4]) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

// cross-GPU barrier of CTA `blockIdx.x` (all ranks launch the same grid): phase 0 before the data phase, 1 after it
__device__ __forceinline__ void nvls_barrier(const NvlsArgs& a, uint32_t tag, int phase) {
  __syncthreads();
  if (threadIdx.x < a.world) {
    __threadfence_system();   // this CTA's prior writes (and, through the CTA barrier, its threads') are visible system-wide
    uint32_t* dst = a.sig[0];
#pragma unroll
    for (int d = 1; d < SYMM_MAX_RANKS; ++d)
      if (d == (int)threadIdx.x) dst = a.sig[d];
    const size_t slot = ((size_t)phase * NVLS_MAX_CTAS + blockIdx.x) * a.world;
    st_release_sys(dst + slot + a.rank, tag);
    const uint32_t* mine = a.sig[0];
#pragma unroll
    for (int d = 1; d < SYMM_MAX_RANKS; ++d)
      if (d == a.rank) mine = a.sig[d];
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + slot + threadIdx.x) != tag) {
      if (clock64() - t0 > 8000000000LL) {
        printf("nvls barrier: rank %d cta %d phase %d timed out waiting for rank %d (tag %u, seen %u)\n", a.rank, (int)blockIdx.x, phase,
               (int)threadIdx.x, tag, ld_acquire_sys(mine + slot + threadIdx.x));
        __trap();
      }
    }
  }
  __syncthreads();
}

// MODE 0 all-reduce, 1 reduce-scatter, 2 all-gather
template <int MODE>
__global__ void __launch_bounds__(NVLS_THREADS) nvls_kernel(const NvlsArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tag = ll_tag(a.step, a.call);
  const int vpr = a.row_elems >> 3;                         // 16-byte vectors per row
  const int my_rows = a.rows_per_seg / a.world;             // rows of one segment owned by a rank
  const long long seg_vec = (long long)my_rows * vpr;       // my vectors inside one segment
  const long long my_vec = seg_vec * a.segs;
  const long long nvec = (long long)a.segs * a.rows_per_seg * vpr;
  const long long tstride = (long long)gridDim.x * NVLS_THREADS;
  const long long t0 = (long long)blockIdx.x * NVLS_THREADS + threadIdx.x;
  // local vector index (inside my slice) -> vector index in the whole buffer
  auto gvec = [&](long long lv) -> long long {
    const long long seg = lv / seg_vec, within = lv % seg_vec;
    return (seg * a.rows_per_seg + (long long)a.rank * my_rows) * vpr + within;
  };
  nvls_barrier(a, tag, 0);     // all-reduce / reduce-scatter: every rank's partial sums are in its copy;
                               // all-gather: every rank has stopped using the previous contents of the buffer
  if (MODE == 2) {
    // all-gather: my slice is already in MY copy; multicast it into everybody's
    for (long long lv = t0; lv < my_vec; lv += tstride) {
      const long long v = gvec(lv);
      const uint4 q = *reinterpret_cast<const uint4*>(a.local + v * 16);
      const uint32_t r[4] = {q.x, q.y, q.z, q.w};
      mm_st_16B(a.mc + v * 16, r);
    }
    nvls_barrier(a, tag, 1);   // everyone's slice has landed everywhere
    return;
  }
  // 4 in-switch reductions in flight per thread (each is a round trip through the NVSwitch)
  for (long long lv0 = t0; lv0 < my_vec; lv0 += tstride * 4) {
    uint32_t r[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long lv = lv0 + j * tstride;
      if (lv < my_vec) mm_ld_reduce_bf16x8(a.mc + gvec(lv) * 16, r[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long lv = lv0 + j * tstride;
      if (lv >= my_vec) continue;
      const long long v = gvec(lv);
      if (MODE == 0) {
        mm_st_16B(a.mc + v * 16, r[j]);             // broadcast the reduced slice into every copy
      } else {
        // reduce-scatter: my slice -> private output rows [segs * my_rows][row_elems] (+ residual of the same shape)
        const long long o = lv * 8;
        if (a.residual != nullptr) {
          const uint4 q = ldg_act(a.residual + o);
          r[j][0] = pack_bf16(bf16lo(r[j][0]) + bf16lo(q.x), bf16hi(r[j][0]) + bf16hi(q.x));
          r[j][1] = pack_bf16(bf16lo(r[j][1]) + bf16lo(q.y), bf16hi(r[j][1]) + bf16hi(q.y));
          r[j][2] = pack_bf16(bf16lo(r[j][2]) + bf16lo(q.z), bf16hi(r[j][2]) + bf16hi(q.z));
          r[j][3] = pack_bf16(bf16lo(r[j][3]) + bf16lo(q.w), bf16hi(r[j][3]) + bf16hi(q.w));
        }
        *reinterpret_cast<uint4*>(a.out + o) = make_uint4(r[j][0], r[j][1], r[j][2], r[j][3]);
      }
    }
  }
  nvls_barrier(a, tag, 1);     // all slices reduced (and broadcast): the buffer may be read / overwritten
  if (MODE == 0 && a.out != nullptr) {
    // private output = reduced buffer (+ residual): every rank reads its own copy
    for (long long v = t0; v < nvec; v += tstride) {
      uint4 q = *reinterpret_cast<const uint4*>(a.local + v * 16);
      if (a.residual != nullptr) {
        const uint4 s = ldg_act(a.residual + v * 8);
        q.x = pack_bf16(bf16lo(q.x) + bf16lo(s.x), bf16hi(q.x) + bf16hi(s.x));
        q.y = pack_bf16(bf16lo(q.y) + bf16lo(s.y), bf16hi(q.y) + bf16hi(s.y));
        q.z = pack_bf16(bf16lo(q.z) + bf16lo(s.z), bf16hi(q.z) + bf16hi(s.z));
        q.w = pack_bf16(bf16lo(q.w) + bf16lo(s.w), bf16hi(q.w) + bf16hi(s.w));
      }
      *reinterpret_cast<uint4*>(a.out + v * 8) = q;
    }
  }
}

void nvls_collective_launch(int mode, const long long* sig_ptrs, const void* step, int rank, int world, int call, void* mc, void* local,
                            const void* residual, void* out, int segs, int rows_per_seg, int row_elems, cudaStream_t stream) {
  if (world < 2 || world > SYMM_MAX_RANKS) throw std::runtime_error("nvls: world must be 2..8");
  if (row_elems % 8 != 0 || rows_per_seg % world != 0) throw std::runtime_error("nvls: rows must split evenly over the ranks, 8-element rows");
  NvlsArgs a{};
  for (int i = 0; i < world; ++i) a.sig[i] = reinterpret_cast<uint32_t*>(sig_ptrs[i]);
  a.step = reinterpret_cast<const uint32_t*>(step);
  a.rank = rank; a.world = world; a.call = call;
  a.mc = reinterpret_cast<uint8_t*>(mc);
  a.local = reinterpret_cast<uint8_t*>(local);
  a.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.segs = segs; a.rows_per_seg = rows_per_seg; a.row_elems = row_elems;
  const long long my_vec = (long long)segs * (rows_per_seg / world) * (row_elems / 8);
  int grid = (int)std::min<long long>(NVLS_MAX_CTAS, std::max<long long>(1, (my_vec + NVLS_THREADS - 1) / NVLS_THREADS));
  if (mode == 0) launch_pdl(nvls_kernel<0>, dim3(grid), dim3(NVLS_THREADS), 0, stream, a);
  else if (mode == 1) launch_pdl(nvls_kernel<1>, dim3(grid), dim3(NVLS_THREADS), 0, stream, a);
  else launch_pdl(nvls_kernel<2>, dim3(grid), dim3(NVLS_THREADS), 0, stream, a);
}

}  // namespace nxdi
