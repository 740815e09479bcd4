// Raw-pointer launcher API between the CUDA translation units (compiled by nvcc, no torch headers) and
// bindings.cpp (compiled by the host compiler against torch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace nxdi {

constexpr int SYMM_MAX_RANKS = 8;
constexpr int SYMM_MAX_TILES = 1024;
constexpr int GEMV_MAX_T = 8;

struct SymmArgs {
  float* recv[SYMM_MAX_RANKS];  // peer-mapped LL receive buffers  [2 parity][world src][8 tok][n_max] x {fp32 value, u32 tag}
  const uint32_t* step;         // device-side step counter (bumped once per forward); tag = (step << 8 | call) + 1
  int rank, world, parity, n_max, call;
};

struct GemvParams {
  const void* x;         // [T, K] bf16
  const void* w;         // [N, K] bf16 (or int8 / fp8 with `scale`)
  const void* bias;      // [N] bf16 or null
  const void* norm_w;    // [K] bf16 or null (fused RMSNorm of x)
  const void* residual;  // [T, N] bf16 or null
  const float* scale;    // [N] (or [1]) fp32 dequant scale of 8-bit weights, or null
  int scale_n;           // numel of `scale`
  void* y;               // [T, N_out] bf16
  int T, N, K, ldx, ldy;
  float eps, norm_offset;
  int act;        // 0 none, 1 silu*up, 2 gelu_tanh*up, 3 gelu*up
  int x_in_smem;  // 0: read (already normalised) x through L1 from global
  int wdtype;     // 0 bf16, 1 int8, 2 fp8 e4m3
  SymmArgs symm;
};

size_t gemv_smem_bytes(int T, int K, bool x_in_smem);
void gemv_launch(const GemvParams& p, int mode, cudaStream_t stream);  // mode 0 plain, 1 fused all-reduce
// v2: bulk-async pipelined, stream-K balanced (gemv2.cu)
bool gemv2_supported(int T, int K, int wt = 0);   // wt: 0 bf16 weights, 1 int8, 2 fp8-e4m3 (weight-only)
int gemv2_grid(int N, int K, bool glu, int wt = 0);
int gemv2_pmax(int N, int K, bool glu, int wt = 0);
int gemv2_ntiles(int N, bool glu);
void gemv2_plan(int N, int K, bool glu, int* rows8, int* whole, int* grid, int* pmax);
int pick_nsplit(int B, int Hkv, int S_hint);
// debug timeline (tools/prof_decode.py): every decode kernel launched while a buffer is set claims the next slot of 148 x 8 u64
void gemv2_set_prof(unsigned long long* base, long long n_launches);
long long prof_next_slot();
long long prof_count();
unsigned long long* prof_slot_ptr(long long slot);
void gemv2_launch(const GemvParams& p, int mode, float* ws_part, unsigned* tickets, cudaStream_t stream);
// tcgen05 / TMEM / TMA GEMM (gemm_tcgen05.cu): C[M,N_out] = epi(A[M,K] · B[N,K]^T)
// fused GEMM -> in-switch reduce-scatter (see GemmParams in gemm_tcgen05.cu)
struct GemmRsArgs {
  long long flag_ptrs[SYMM_MAX_RANKS];   // peer-mapped [max_tiles][world] u32
  const void* mc;                        // multicast address of the staging buffer `c`
  void* out;                             // private output [M / world, N]
  const void* residual;                  // [M / world, N] or null
  const void* step;
  int call, rank, world, rows_per_seg, max_tiles;
  // all-reduce flavour: multicast address of the symmetric output [M, N] (null: reduce-scatter), completion flags, CTA counter
  const void* bcast_mc;
  long long done_ptrs[SYMM_MAX_RANKS];
  void* cta_counter;
};
void gemm_tcgen05_launch(const void* a, int lda, const void* b, const void* bias, const void* residual, void* c, int ldc, int M,
                         int N, int K, int act, cudaStream_t stream, const GemmRsArgs* rs = nullptr);
void rmsnorm_launch(const void* x, const void* res_in, const void* w, void* y, void* res_out, int rows, int H, float eps,
                    float offset, cudaStream_t stream);

void rope_kv_append_launch(const void* qkv, const float* cos, const float* sin, void* q_out, void* k_cache, void* v_cache,
                           const int* lines, const int* positions, const void* q_norm, const void* k_norm, float eps, int B,
                           int T, int nq, int nkv, int D, int L, int S, cudaStream_t stream, void* k_out = nullptr, void* v_out = nullptr);
void kv_append_launch(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* lines,
                      const int* positions, int B, int T, int H, int row_bytes, int L, int S, cudaStream_t stream);
void paged_kv_append_launch(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* slots, int ntok,
                            int tok_bytes, int n_slots, cudaStream_t stream);

// vocabulary-sharded arg-max: per-row {value, global index} exchanged over NVLink LL slots inside the kernel
struct ArgmaxSymm {
  float* slots[SYMM_MAX_RANKS];  // peer-mapped  [2 parity][rows_max][world src][2] x {payload, u32 tag}
  const uint32_t* step;
  int rank, world, parity, call, rows_max;
};
void argmax_launch(const void* logits, int dtype /*0 f32, 1 bf16*/, int64_t* out, float* ws_val, int* ws_idx,
                   unsigned* tickets, int B, int V, int ld, int nsplit, const ArgmaxSymm* symm, cudaStream_t stream);
void topk_sample_launch(const void* logits, int dtype, const int* top_k, const float* top_p, const float* temperature,
                        const float* rand, int64_t* out, int B, int V, int ld, int K, cudaStream_t stream);

// weight-only int8 / fp8-e4m3 skinny GEMM (qgemv.cu): wdtype 1 = int8, 2 = fp8 e4m3; scale fp32 [N] or [1]
void qgemv_launch(const void* x, const void* w, const float* scale, int scale_n, const void* bias, const void* norm_w,
                  const void* residual, void* y, int T, int N, int K, int ldx, int ldy, int act, int wdtype, float eps,
                  float norm_offset, int n_sms, cudaStream_t stream);

// MoE decode: routed experts for T <= 8 tokens as two batched streaming launches (moe_decode.cu)
void moe_decode_launch(const void* x, const void* w_gate_up, const void* w_down, const float* topk_w, const int* topk_i, void* u,
                       float* y_acc, int T, int topk, int H, int I, int E, int expert_offset, int n_sms, cudaStream_t stream);

struct AttnDecodeParams {
  const void* q;        // [B, T, Hq, D] bf16
  const void* k_cache;  // contiguous [L, Hkv, S, D] or paged [nblk, bs, Hkv, D]
  const void* v_cache;
  void* out;            // [B, T, Hq, D] bf16
  const int* lines;     // [B] cache line per row (contiguous) or null
  const int* positions;  // [B, T] absolute position of each active token (attends to keys <= position)
  const int* block_table;  // paged: [B, max_blocks] or null
  const float* sinks;      // [Hq] or null
  float* ws_o;             // split-KV workspace: [B*Hkv, nsplit, R, D] fp32
  float* ws_ml;            // [B*Hkv, nsplit, R, 2]
  unsigned* tickets;       // [B*Hkv]
  int B, T, Hq, Hkv, D, S, L, nsplit, window, block_size, max_blocks;
  float scale;
  // fused prologue (contiguous decode): q unused; q/k/v taken from the QKV projection output
  const void* qkv;        // [B, T, (Hq + 2 Hkv) * D] bf16 or null
  const float* cos;       // [B, T, D/2]
  const float* sin;
  const void* q_norm;     // [D] bf16 or null (per-head RMSNorm before RoPE)
  const void* k_norm;
  const int* write_pos;   // [B, T] cache slot per active token
  float norm_eps;
};
void attention_decode_launch(const AttnDecodeParams& p, cudaStream_t stream);

struct AttnPrefillParams {
  const void* q;  // [B, T, Hq, D]
  const void* k;  // [B, T, Hkv, D]
  const void* v;
  void* out;      // [B, T, Hq, D]
  const float* sinks;
  int B, T, Hq, Hkv, D, window;
  float scale;
  int causal;
};
void attention_prefill_launch(const AttnPrefillParams& p, cudaStream_t stream);
// tcgen05 / TMEM / TMA flash attention (attention_tc.cu), head_dim 128
void attention_prefill_tc_launch(const void* q, const void* k, const void* v, void* out, const float* sinks, int B, int T, int Hq, int Hkv,
                                 float scale, int causal, int window, float softcap, cudaStream_t stream);

// W8A8 fp8 path: dynamic per-token activation quantisation (quant.cu) + tcgen05 kind::f8f6f4 GEMM (gemm_tcgen05.cu)
void rmsnorm_quant_launch(const void* x, const void* gamma, void* q, float* scale, int rows, int H, float eps, float offset, float clamp,
                          cudaStream_t stream);
void dequant_bf16_launch(const void* w, int wdtype, const float* scale, int scale_n, void* out, int N, int K, cudaStream_t stream);
void gemm_fp8_launch(const void* a, int lda, const void* b, const float* a_scale, const float* w_scale, int w_scale_n, const void* bias,
                     const void* residual, void* c, int ldc, int M, int N, int K, int act, cudaStream_t stream);

// mixture-of-experts prefill: device-side permutation (moe_grouped.cu) around the grouped tcgen05 GEMM (gemm_tcgen05.cu)
void moe_plan_launch(const int* topk_i, int entries, int expert_offset, int e_local, int R, int* pos, int* row_entry, int* tile_expert,
                     cudaStream_t stream);
void moe_gather_launch(const void* x, const float* topk_w, const int* row_entry, const int* tile_expert, void* xp, int R, int H, int k,
                       int scale_input, cudaStream_t stream);
void moe_combine_launch(const void* y, const float* topk_w, const int* pos, void* out, int N, int H, int k, int scale_input,
                        cudaStream_t stream);
void gemm_grouped_launch(const void* a, const void* b, const void* bias, void* c, int R, int N, int K, int n_experts, int act,
                         const int* tile_expert, const float* a_scale, const float* w_scale, cudaStream_t stream);

// symmetric heap (symm_heap.cpp) and NVLS collectives (nvls.cu)
long long symm_heap_create(long long bytes, int device, int world, int rank);
long long symm_heap_size(long long h);
long long symm_heap_local_va(long long h);
int symm_heap_export_fd(long long h);
long long symm_heap_import_peer(long long h, int peer, int fd);
bool symm_heap_multicast_supported(int device);
int symm_heap_mc_create(long long h, int world);
void symm_heap_mc_import(long long h, int fd);
void symm_heap_mc_add_device(long long h);
long long symm_heap_mc_bind_map(long long h);
long long symm_heap_mc_va(long long h);
void symm_heap_destroy(long long h);
constexpr int NVLS_SIG_WORDS = 2 * 64 * SYMM_MAX_RANKS;   // [2 phase][64 CTAs][world] u32 signal words per heap
void nvls_collective_launch(int mode, const long long* sig_ptrs, const void* step, int rank, int world, int call, void* mc, void* local,
                            const void* residual, void* out, int segs, int rows_per_seg, int row_elems, cudaStream_t stream);

// persistent decode-step kernel (decode_step.cu): all decoder layers of one decode step in one launch
long long dstep_new(int T);
void dstep_set_symm(long long h, const std::vector<long long>& recv_ptrs, const void* step, int rank, int n_max);
void dstep_add_gemv(long long h, const void* w, int N, int K, const void* x, int ldx, const void* bias, const void* norm_w, float eps,
                    float norm_offset, int act, const void* residual, void* y, int ldy, bool allreduce);
void dstep_add_attn(long long h, const void* qkv, void* out, void* k_cache, void* v_cache, const void* q_norm, const void* k_norm,
                    float norm_eps, int B, int T, int Hq, int Hkv, int D, int S, int L, float scale, int window, const float* sinks,
                    int nsplit);
void dstep_finalize(long long h);
int dstep_num_allreduce(long long h);
void dstep_free(long long h);
void dstep_launch(long long h, const int* positions, const int* write_pos, const int* lines, const float* cos, const float* sin,
                  int call_base, int parity_base, cudaStream_t stream);

}  // namespace nxdi
