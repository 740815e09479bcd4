// GEMV v2: TMA (3-D tensor map, UTMALDG) mbarrier-pipelined weight-streaming skinny GEMM for decode
//          Y[T<=8, N] = f( rmsnorm(X)[T,K] · W[N,K]^T )      (+ fused one-shot all-reduce over NVLink)
//
// Design (round 2):
//   * ONE elected producer thread streams W with TMA tensor loads (cp.async.bulk.tensor.3d, completion on an
//     mbarrier): W[N,K] is described as a 3-D tensor {64 k, K/64 groups, N rows}; one instruction fetches the box
//     {64, 2, 16} = 16 weight rows x 128 k = 4 KB into a 128B-swizzled stage.
//   * CO-RESIDENCY: the CTA is sized to <= ~108 KB of shared memory so that TWO consecutive decode kernels fit on an SM.
//     Under programmatic dependent launch the NEXT kernel's producer then fills its ring (weights never depend on the
//     previous kernel) while THIS kernel is still computing: HBM no longer idles for the ~5 us launch / wait / x-prologue /
//     flush window between dependent GEMVs, which was what separated 3.17 ms/step from the 2.28 ms streaming floor at TP1 and
//     dominated the step at TP4/TP8 (profiles/decode_overheads_r1.md).
//   * 8 consumer warps; a stage always belongs to warp (stage & 7), so a warp meets its stages lap after lap in order and the
//     mbarrier phase parity is unambiguous for ANY ring depth >= 8.  A-fragments of mma.sync m16n8k16 come straight from the
//     swizzled stage (8-lane LDS.128 phases are conflict free); the token fragments of X (gamma folded in, bf16) sit in shared
//     memory; 1/rms is applied to the fp32 accumulators in the epilogue (single pass over x).
//   * stream-K for large weights (equal bytes per SM, ticketed deterministic fix-up), whole 16-row tiles per CTA for small ones.
//   * fused all-reduce (MODE 1): LL protocol — every partial travels as an 8-byte {value, tag} store straight into each peer's
//     receive slot over NVLink; the receiver polls all sources of an element IN PARALLEL (one round trip, not `world`).  The
//     tag is an epoch = (device-side step counter, call index in the step): a skipped / repeated collective can no longer be
//     mistaken for the expected one (it times out and traps instead), and slots need no reset.  Slots are double-buffered by
//     call parity (a rank can be at most one collective ahead of its slowest peer).
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "gemv2_body.cuh"

namespace nxdi {

struct Gemv2Params {
  CUtensorMap tmap;
  G2Phase ph;       // ph.tmap is patched to &tmap inside the kernel (parameter space)
  G2Symm symm;
  int n_stages;
  int max_inflight;
};

template <bool GLU, int MODE, int WT>
__global__ void __launch_bounds__(G2_THREADS, 1) gemv2_kernel(const __grid_constant__ Gemv2Params pp) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  G2Phase ph = pp.ph;
  ph.tmap = &pp.tmap;
  ph.grid = gridDim.x;
  // shared memory carve-up (stage_base 1024-aligned for the 128B swizzle):
  //   [stages][16 KB] | xs[T][2*Kp+64] | red[8][128] f32 | rstd[64] f32 | barriers        Kp = K rounded up to 512
  __shared__ int s_flag;
  G2Smem sm;
  sm.NS = pp.n_stages;
  const int Kp = (ph.K + g2_kc(WT) - 1) / g2_kc(WT) * g2_kc(WT);
  sm.stage_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: keeps the address space
  sm.xs = sm.stage_base + (size_t)sm.NS * G2_STAGE_BYTES;
  sm.red = reinterpret_cast<float*>(sm.xs + (size_t)ph.T * (Kp * 2 + 64));
  sm.rstd_s = sm.red + G2_CONSUMER_WARPS * 128;
  sm.full_bar = reinterpret_cast<uint64_t*>(sm.rstd_s + 64);
  sm.empty_bar = sm.full_bar + G2_MAX_STAGES;
  sm.x_bar = sm.empty_bar + G2_MAX_STAGES;
  sm.s_flag = &s_flag;
  sm.max_inflight = pp.max_inflight;
  if (tid == 0) {
    for (int s = 0; s < sm.NS; ++s) {
      mbar_init(&sm.full_bar[s], 1);
      mbar_init(&sm.empty_bar[s], G2_CONSUMER_WARPS);
    }
    mbar_init(sm.x_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == G2_CONSUMER_WARPS) {
    // producer warp.  The dependent kernel may start right away: it only prefetches ITS weights until this grid has completed.
    pdl_launch_dependents();
    if (lane == 0) {
      int stage = 0, issued = 0;
      uint32_t phb = 0;
      g2_produce(ph, sm, blockIdx.x, stage, phb, issued);
    }
    return;
  }
  int stage = 0;
  uint32_t lap = 0, xph = 0;
  g2_consume<GLU, MODE, WT>(ph, pp.symm, sm, blockIdx.x, tid, stage, lap, xph, [] { pdl_wait(); });
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

static size_t g2_fixed_smem(int T, int K, int wt = 0) {
  const int Kp = (K + g2_kc(wt) - 1) / g2_kc(wt) * g2_kc(wt);
  return (size_t)T * (Kp * 2 + 64) + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + (2 * G2_MAX_STAGES + 1) * sizeof(uint64_t) +
         128 + 1024;  // + worst-case 1 KB alignment slack
}

bool gemv2_supported(int T, int K, int wt) {
  return K % (wt ? 128 : 64) == 0 && g2_fixed_smem(T, K, wt) + 3 * G2_STAGE_BYTES <= G2_SMEM_BUDGET;
}

// Small weights: whole tiles per CTA (no cross-CTA fix-up).  Large weights: stream-K (perfect byte balance matters more).
static bool g2_rows8(int N, bool glu) { return !glu && N % 8 == 0 && (N + 15) / 16 < 100; }
static int g2_ntiles(int N, bool glu) { return glu ? ((N / 2) + 7) / 8 : (g2_rows8(N, glu) ? (N + 7) / 8 : (N + 15) / 16); }
static double g2_tile_bytes(int N, int K, bool glu, int wt) { return (g2_rows8(N, glu) ? 8.0 : 16.0) * K * (wt ? 1 : 2); }

static bool g2_whole_tiles(int N, int K, bool glu, int wt = 0) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NXDI_B200_GEMV_WHOLE_TILES");
    mode = e ? atoi(e) : 2;   // 0 never, 1 always, 2 heuristic
  }
  if (mode != 2) return mode == 1;
  const int n_tiles = g2_ntiles(N, glu);
  const int sms = g2_num_sms();
  const int active = std::min(sms, n_tiles);
  const int per = (n_tiles + active - 1) / active;            // tiles of the busiest CTA
  const double tile_bytes = g2_tile_bytes(N, K, glu, wt);
  // one SM pulls at most ~100 GB/s; all of them together ~6 TB/s (bytes per microsecond below)
  const double bw_sm = std::min(100e3, 6.0e6 / active);
  const double t_whole = per * tile_bytes / bw_sm;
  // stream-K: perfectly balanced bytes + the fix-up (partials, ticket, re-read: ~3.5-4 us; tools/bench_gemv_fixed.py:
  // 4096x4096 11.1 -> 6.8 us, 1536x4096 10.4 -> 4.6 us once it is gone; Llama-8B decode 3.67 -> 3.19 ms/step)
  const double t_streamk = (double)N * K * (wt ? 1 : 2) / 6.0e6 + 4.0;
  return t_whole <= t_streamk + 0.5;
}

int gemv2_grid(int N, int K, bool glu, int wt) {
  const int n_tiles = g2_ntiles(N, glu);
  if (g2_whole_tiles(N, K, glu, wt)) return std::min(g2_num_sms(), n_tiles);
  const long long U = (long long)n_tiles * ((K + g2_kc(wt) - 1) / g2_kc(wt));
  return (int)std::min<long long>(g2_num_sms(), std::max<long long>(U / 2, 1));   // >= 32 KB of weights per CTA
}

int gemv2_ntiles(int N, bool glu) { return g2_ntiles(N, glu); }
int g2_num_sms_public() { return g2_num_sms(); }
int g2_max_inflight() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NXDI_B200_GEMV_INFLIGHT");
    v = e ? atoi(e) : 6;   // measured (profiles/decode_r2.md): TP1 2.94 ms/step uncapped -> 2.82 with 6 stages (96 KB) in flight
  }
  return v;
}
void gemv2_plan(int N, int K, bool glu, int* rows8, int* whole, int* grid, int* pmax) {
  *rows8 = g2_rows8(N, glu) ? 1 : 0;
  *whole = g2_whole_tiles(N, K, glu) ? 1 : 0;
  *grid = gemv2_grid(N, K, glu);
  *pmax = gemv2_grid(N, K, glu) / g2_ntiles(N, glu) + 3;
}

int gemv2_pmax(int N, int K, bool glu, int wt) {
  const int n_tiles = g2_ntiles(N, glu);
  return gemv2_grid(N, K, glu, wt) / n_tiles + 3;
}

template <bool GLU, int MODE, int WT>
static void launch_gemv2(const Gemv2Params& pp, cudaStream_t stream) {
  auto kern = gemv2_kernel<GLU, MODE, WT>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BUDGET);
    configured = true;
  }
  const size_t smem = g2_fixed_smem(pp.ph.T, pp.ph.K, WT) + (size_t)pp.n_stages * G2_STAGE_BYTES;
  const int grid = gemv2_grid(pp.ph.N, pp.ph.K, GLU, WT);
  launch_pdl(kern, dim3(grid), dim3(G2_THREADS), smem, stream, pp);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// W [N, K] bf16 row-major as a 3-D tensor {64 k, N rows, K/64 groups} (strides 2 B, K*2 B, 128 B); box {64, rows, 8}:
// one instruction fetches rows x 512 k, and the box lands in shared memory as [k group][row][64 k] — 128-byte lines whose
// index & 7 is the row & 7, so the eight rows an LDS.128 phase touches sit in eight different swizzle positions.
// (8-bit weights: {128 k, N rows, K/128 groups} of bytes — the same 128-byte lines, twice the k per stage)
void make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K, int box_rows, int wt) {
  const int line = wt ? 128 : 64, es = wt ? 1 : 2;
  cuuint64_t gdim[3] = {(cuuint64_t)line, (cuuint64_t)N, (cuuint64_t)(K / line)};
  cuuint64_t gstride[2] = {(cuuint64_t)K * es, 128};
  cuuint32_t box[3] = {(cuuint32_t)line, (cuuint32_t)box_rows, (cuuint32_t)G2_KG};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_tiled()(tm, wt ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
}

// Ring depth.  Default budget G2_SMEM_CORES keeps the CTA small enough for TWO decode kernels per SM (see the header);
// never deeper than the busiest CTA's unit count (a 2-tile CTA does not need 22 stages); at least 8 stages (one per warp).
// If even 8 stages do not fit the co-residency budget (wide K x many tokens) the kernel takes what it needs up to 224 KB.
static int g2_pick_stages(const GemvParams& p, bool glu, bool whole) {
  static int budget_kb = -1;
  if (budget_kb < 0) {
    const char* e = getenv("NXDI_B200_GEMV_SMEM_KB");
    budget_kb = e ? atoi(e) : G2_SMEM_CORES / 1024;
  }
  const int wt = p.scale != nullptr ? p.wdtype : 0;
  const size_t fixed = g2_fixed_smem(p.T, p.K, wt);
  const size_t budget = std::min<size_t>((size_t)budget_kb * 1024, G2_SMEM_BUDGET);
  int ns = budget > fixed ? (int)((budget - fixed) / G2_STAGE_BYTES) : 0;
  const int n_tiles = g2_ntiles(p.N, glu);
  const int n_chunks = (p.K + g2_kc(wt) - 1) / g2_kc(wt);
  const int G = gemv2_grid(p.N, p.K, glu, wt);
  const long long per_cta = whole ? (long long)((n_tiles + G - 1) / G) * n_chunks : ((long long)n_tiles * n_chunks + G - 1) / G;
  ns = (int)std::min<long long>(ns, per_cta);
  ns = std::min(ns, G2_MAX_STAGES);
  if (ns < 3) {   // wide activations: give up co-residency, take what the SM has
    ns = (int)std::min<long long>((G2_SMEM_BUDGET - fixed) / G2_STAGE_BYTES, std::max<long long>(per_cta, 3));
    ns = std::min(ns, G2_MAX_STAGES);
  }
  return std::max(ns, 2);
}

static unsigned long long* g_prof_base = nullptr;
static long long g_prof_cap = 0, g_prof_next = 0;
void gemv2_set_prof(unsigned long long* base, long long n_launches) { g_prof_base = base; g_prof_cap = n_launches; g_prof_next = 0; }
long long prof_count() { return g_prof_next; }
long long prof_next_slot() { return (g_prof_base && g_prof_next < g_prof_cap) ? g_prof_next++ : -1; }
unsigned long long* prof_slot_ptr(long long slot) { return slot < 0 ? nullptr : g_prof_base + (size_t)slot * 148 * 8; }

void gemv2_launch(const GemvParams& p, int mode, float* ws_part, unsigned* tickets, cudaStream_t stream) {
  Gemv2Params pp{};
  G2Phase& ph = pp.ph;
  ph.x = p.x; ph.bias = p.bias; ph.norm_w = p.norm_w; ph.residual = p.residual; ph.y = p.y;
  ph.T = p.T; ph.N = p.N; ph.K = p.K; ph.ldx = p.ldx; ph.ldy = p.ldy; ph.eps = p.eps; ph.norm_offset = p.norm_offset; ph.act = p.act;
  ph.parity = p.symm.parity; ph.call = p.symm.call;
  for (int i = 0; i < SYMM_MAX_RANKS; ++i) pp.symm.recv[i] = p.symm.recv[i];
  pp.symm.step = p.symm.step; pp.symm.rank = p.symm.rank; pp.symm.world = p.symm.world; pp.symm.n_max = p.symm.n_max;
  ph.prof = prof_slot_ptr(prof_next_slot());
  const bool glu = p.act != 0;
  const int wt = p.scale != nullptr ? p.wdtype : 0;
  ph.wt = wt; ph.wscale = p.scale; ph.wscale_n = p.scale_n;
  ph.rows8 = g2_rows8(p.N, glu) ? 1 : 0;
  make_weight_tmap(&pp.tmap, p.w, p.N, p.K, (glu || ph.rows8) ? 8 : 16, wt);
  ph.ws_part = ws_part;
  ph.tickets = tickets;
  ph.p_max = gemv2_pmax(p.N, p.K, glu, wt);
  ph.whole_tiles = g2_whole_tiles(p.N, p.K, glu, wt) ? 1 : 0;
  pp.n_stages = g2_pick_stages(p, glu, ph.whole_tiles != 0);
  pp.max_inflight = g2_max_inflight() < pp.n_stages ? g2_max_inflight() : 0;
#define G2_LAUNCH(WT)                                      \
  if (mode == 1) launch_gemv2<false, 1, WT>(pp, stream);   \
  else if (glu) launch_gemv2<true, 0, WT>(pp, stream);     \
  else launch_gemv2<false, 0, WT>(pp, stream);
  if (wt == 0) { G2_LAUNCH(0) }
  else if (wt == 1) { G2_LAUNCH(1) }
  else { G2_LAUNCH(2) }
#undef G2_LAUNCH
}

}  // namespace nxdi
