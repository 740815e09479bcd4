// Launchers of the flash-attention kernels; the kernel body lives in attention_body.cuh (shared with decode_step.cu).
#include "attention_body.cuh"

namespace nxdi {

template <int D, int MODE, int ATT_STAGES>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnArgs p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  attention_body<D, MODE, ATT_STAGES>(p, smem_raw, threadIdx.x, blockIdx.x, blockIdx.y, gridDim.x, 0, true, [] { pdl_wait(); });
}

template <int D, int MODE, int ATT_STAGES>
static void launch_attn(const AttnArgs& a, dim3 grid, cudaStream_t stream) {
  auto kern = attention_kernel<D, MODE, ATT_STAGES>;
  // + this step's K/V rows (fused decode: [2][T][D], patched into the resident tiles)
  const size_t extra = (MODE == ATTN_DECODE && a.qkv != nullptr) ? (size_t)2 * a.T * D * sizeof(__nv_bfloat16) : 0;
  const size_t smem = (size_t)(64 * D + 2 * ATT_STAGES * 64 * D) * sizeof(__nv_bfloat16) + extra;
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  launch_pdl(kern, grid, dim3(ATT_THREADS), smem, stream, a);
}

static bool grid_fits_prof(int ctas) { return ctas <= 148; }

void attention_decode_launch(const AttnDecodeParams& p, cudaStream_t stream) {
  AttnArgs a{};
  a.q = reinterpret_cast<const __nv_bfloat16*>(p.q);
  a.k = reinterpret_cast<const __nv_bfloat16*>(p.k_cache);
  a.v = reinterpret_cast<const __nv_bfloat16*>(p.v_cache);
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.lines = p.lines;
  a.positions = p.positions;
  a.block_table = p.block_table;
  a.sinks = p.sinks;
  a.ws_o = p.ws_o;
  a.ws_ml = p.ws_ml;
  a.tickets = p.tickets;
  a.B = p.B; a.T = p.T; a.Hq = p.Hq; a.Hkv = p.Hkv; a.S = p.S; a.L = p.L; a.nsplit = p.nsplit; a.window = p.window;
  a.block_size = p.block_size; a.max_blocks = p.max_blocks;
  a.scale_log2 = p.scale * kLog2e;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  a.cos = p.cos; a.sin = p.sin;
  a.q_norm = reinterpret_cast<const __nv_bfloat16*>(p.q_norm);
  a.k_norm = reinterpret_cast<const __nv_bfloat16*>(p.k_norm);
  a.write_pos = p.write_pos;
  a.k_w = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(p.k_cache));
  a.v_w = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(p.v_cache));
  a.norm_eps = p.norm_eps;
  a.prof = (grid_fits_prof(p.B * p.Hkv * p.nsplit)) ? prof_slot_ptr(prof_next_slot()) : nullptr;
  if (p.T * (p.Hq / p.Hkv) > 64) throw std::runtime_error("attention_decode: T * group size must be <= 64");
  dim3 grid(p.B * p.Hkv, p.nsplit);
  const bool paged = p.block_table != nullptr;
  const bool deep = p.nsplit == 1;  // short context, single split: latency bound -> deep prefetch
  if (p.D == 128) {
    if (paged) { if (deep) launch_attn<128, ATTN_PAGED, 3>(a, grid, stream); else launch_attn<128, ATTN_PAGED, 2>(a, grid, stream); }
    else { if (deep) launch_attn<128, ATTN_DECODE, 3>(a, grid, stream); else launch_attn<128, ATTN_DECODE, 2>(a, grid, stream); }
  } else if (p.D == 64) {
    if (paged) { if (deep) launch_attn<64, ATTN_PAGED, 3>(a, grid, stream); else launch_attn<64, ATTN_PAGED, 2>(a, grid, stream); }
    else { if (deep) launch_attn<64, ATTN_DECODE, 3>(a, grid, stream); else launch_attn<64, ATTN_DECODE, 2>(a, grid, stream); }
  } else {
    throw std::runtime_error("attention_decode: head_dim must be 64 or 128");
  }
}

void attention_prefill_launch(const AttnPrefillParams& p, cudaStream_t stream) {
  AttnArgs a{};
  a.q = reinterpret_cast<const __nv_bfloat16*>(p.q);
  a.k = reinterpret_cast<const __nv_bfloat16*>(p.k);
  a.v = reinterpret_cast<const __nv_bfloat16*>(p.v);
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.sinks = p.sinks;
  a.B = p.B; a.T = p.T; a.Hq = p.Hq; a.Hkv = p.Hkv; a.nsplit = 1; a.window = p.window;
  a.causal = p.causal;
  a.scale_log2 = p.scale * kLog2e;
  dim3 grid(p.B * p.Hq * ((p.T + 63) / 64));
  if (p.D == 128) launch_attn<128, ATTN_PREFILL, 2>(a, grid, stream);
  else if (p.D == 64) launch_attn<64, ATTN_PREFILL, 2>(a, grid, stream);
  else throw std::runtime_error("attention_prefill: head_dim must be 64 or 128");
}

}  // namespace nxdi
