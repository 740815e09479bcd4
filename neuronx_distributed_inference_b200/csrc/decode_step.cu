// Persistent decode-step kernel: ALL decoder layers of one decode step in ONE launch.
//
// Why: a decode layer is five dependent kernels (qkv -> attention -> o_proj(+all-reduce) -> gate_up -> down(+all-reduce)).  With
// the weights sharded 8 ways a layer streams 55 MB (8 us of HBM time) but took ~38 us: every kernel boundary costs 5-6 us of
// grid drain + dependency release + activation reload + ring refill (profiles/decode_r2.md, tools/prof_decode.py), and no
// amount of PDL overlap hides it because the next kernel cannot consume before its input exists.  Here one CTA per SM stays
// resident for the whole step:
//   * the TMA producer thread walks the phase list and streams the weights of EVERY projection of EVERY layer back to back
//     through the CTA's mbarrier ring — weights never depend on activations, so HBM keeps streaming across phase boundaries
//     (the ring is ~190 KB per SM = 28 MB chip-wide, ~4 us of HBM time of run-ahead);
//   * the 8 consumer warps run the phases in order; a phase boundary is a device-side counter: the CTAs that produced a
//     phase's output do `red.release.gpu` on it, the consumers of the next phase poll it (`ld.acquire.gpu`) and then fetch
//     the activations from L2 — ~2 us instead of a kernel boundary;
//   * attention is a phase too (the same `attention_body` as the stand-alone kernel, run by warps 0-3 of the first
//     B*Hkv*nsplit CTAs; those CTAs give the tail of their ring to the K/V tiles);
//   * the tensor-parallel all-reduces stay fused in the o_proj / down epilogues (LL stores into every peer over NVLink).
// Reference: the per-layer kernels this replaces are K2-K5 of SURVEY §2.3 (attention_block_tkg, fused QKV, o_proj + reduce,
// MLP) — the reference fuses WITHIN a block; on B200 the step itself is the unit.
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "attention_body.cuh"
#include "gemv2_body.cuh"

namespace nxdi {

enum { PH_GEMV = 0, PH_ATTN = 1 };

struct StepPhase {
  int type;      // PH_GEMV / PH_ATTN
  int expected;  // CTAs that signal the completion of THIS phase
  int n_items;   // PH_ATTN: work items (batch x kv head x split), one per CTA
  int pad_;
  G2Phase g;
  AttnArgs a;    // static part; positions / lines / cos / sin are patched from StepParams
};

struct StepParams {
  const StepPhase* phases;
  unsigned* done;   // [n_phases], zeroed before the launch
  int n_phases;
  int n_stages;     // ring depth (CTAs without attention work)
  int attn_stages;  // ring stages an attention CTA hands to the K/V tiles
  int max_items;    // largest n_items over the attention phases
  int xs_bytes;     // activation staging area (max over the GEMV phases)
  int call_base;    // tag / parity of the first all-reduce of this launch (host-side running counters of the workspace)
  int parity_base;
  int max_inflight;
  G2Symm symm;
  // inputs of this step (fresh tensors every call, so they are launch parameters, not part of the phase table)
  const int* positions;
  const int* write_pos;
  const int* lines;
  const float* cos;
  const float* sin;
  unsigned long long* prof;   // debug timeline: [n_phases][148][8] u64 or null (tools/prof_decode.py)
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread: spin until `n` CTAs have published phase `j` (bounded: a lost signal traps instead of hanging the GPU)
__device__ __forceinline__ void poll_done(const unsigned* done, int j, unsigned n) {
  const long long t0 = clock64();
  while (ld_acquire_gpu_u32(done + j) < n) {
    if (clock64() - t0 > 6000000000LL) {
      printf("decode_step: cta %d timed out waiting for phase %d (%u of %u)\n", (int)blockIdx.x, j, ld_acquire_gpu_u32(done + j), n);
      __trap();
    }
  }
}

template <int D>
__global__ void __launch_bounds__(G2_THREADS, 1) decode_step_kernel(const __grid_constant__ StepParams sp) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const bool attn_cta = c < sp.max_items;
  G2Smem sm;
  sm.NS = attn_cta ? sp.n_stages - sp.attn_stages : sp.n_stages;
  sm.stage_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: keeps the address space
  uint8_t* attn_smem = sm.stage_base + (size_t)sm.NS * G2_STAGE_BYTES;   // only meaningful on attention CTAs
  sm.xs = sm.stage_base + (size_t)sp.n_stages * G2_STAGE_BYTES;
  sm.red = reinterpret_cast<float*>(sm.xs + sp.xs_bytes);
  sm.rstd_s = sm.red + G2_CONSUMER_WARPS * 128;
  sm.full_bar = reinterpret_cast<uint64_t*>(sm.rstd_s + 64);
  sm.empty_bar = sm.full_bar + G2_MAX_STAGES;
  sm.x_bar = sm.empty_bar + G2_MAX_STAGES;
  sm.s_flag = &s_flag;
  sm.max_inflight = sp.max_inflight;
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
    // ===================== producer: the weights of every phase, back to back =====================
    pdl_launch_dependents();
    if (lane == 0) {
      int stage = 0, issued = 0;
      uint32_t phb = 0;
      for (int i = 0; i < sp.n_phases; ++i) {
        const StepPhase& P = sp.phases[i];
        if (P.type == PH_GEMV) g2_produce(P.g, sm, c, stage, phb, issued);
      }
    }
    return;
  }

  // ===================== consumers =====================
  int stage = 0;
  uint32_t lap = 0, xph = 0;
  for (int i = 0; i < sp.n_phases; ++i) {
    const StepPhase& P = sp.phases[i];
    if (P.type == PH_GEMV) {
      G2Phase g = P.g;
      if (g.parity >= 0) {
        g.parity ^= sp.parity_base;
        g.call += sp.call_base;
      }
      g.prof = sp.prof ? sp.prof + (size_t)i * 148 * 8 : nullptr;
      auto wait_dep = [&] {
        if (i == 0) {
          pdl_wait();   // the kernels before this step (embedding, rotary tables)
        } else {
          if (tid == 0) poll_done(sp.done, i - 1, (unsigned)sp.phases[i - 1].expected);
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      };
      if (g.act != 0) g2_consume<true, 0>(g, sp.symm, sm, c, tid, stage, lap, xph, wait_dep);
      else if (g.parity >= 0) g2_consume<false, 1>(g, sp.symm, sm, c, tid, stage, lap, xph, wait_dep);
      else g2_consume<false, 0>(g, sp.symm, sm, c, tid, stage, lap, xph, wait_dep);
      // publish: every store of this CTA for the phase is ordered before the counter update (barrier + release)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0 && c < P.expected) red_release_gpu_add(sp.done + i, 1u);
    } else {
      if (c < P.n_items && warp < 4) {
        AttnArgs a = P.a;
        a.positions = sp.positions;
        a.write_pos = sp.write_pos;
        a.lines = sp.lines;
        a.cos = sp.cos;
        a.sin = sp.sin;
        a.prof = sp.prof ? sp.prof + (size_t)i * 148 * 8 : nullptr;
        const int gdx = a.B * a.Hkv;
        auto wait_dep = [&] {
          if (tid == 0) poll_done(sp.done, i - 1, (unsigned)sp.phases[i - 1].expected);
          asm volatile("bar.sync 2, 128;" ::: "memory");
        };
        attention_body<D, ATTN_DECODE, 2>(a, attn_smem, tid, c % gdx, c / gdx, gdx, 2, false, wait_dep);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (tid == 0) red_release_gpu_add(sp.done + i, 1u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Host side: a step plan = the phase table (+ its tensor maps) in device memory, built once per (model, batch, tokens).
struct StepPlan {
  std::vector<StepPhase> phases;
  std::vector<CUtensorMap> tmaps;
  std::vector<int> tmap_index;   // per phase, -1 for attention
  void* d_phases = nullptr;
  void* d_tmaps = nullptr;
  unsigned* d_done = nullptr;
  float* d_ws = nullptr;         // stream-K partials (shared by the phases: a phase only starts once the previous one has completed)
  unsigned* d_tickets = nullptr;
  float* d_attn_ws = nullptr;
  unsigned* d_attn_tickets = nullptr;
  size_t ws_floats = 0, n_tickets = 0, attn_ws_floats = 0, attn_tickets = 0;
  int T = 0, D = 0, xs_bytes = 0, max_items = 0, n_ar = 0, n_stages = 0, attn_stages = 0;
  size_t smem = 0;
  G2Symm symm{};
  bool finalized = false;
};

static std::vector<StepPlan*>& plans() {
  static std::vector<StepPlan*> v;
  return v;
}
static StepPlan& plan(long long h) {
  if (h < 0 || h >= (long long)plans().size() || plans()[h] == nullptr) throw std::runtime_error("decode_step: bad plan handle");
  return *plans()[h];
}

long long dstep_new(int T) {
  auto* p = new StepPlan();
  p->T = T;
  plans().push_back(p);
  return (long long)plans().size() - 1;
}

void dstep_set_symm(long long h, const std::vector<long long>& recv_ptrs, const void* step, int rank, int n_max) {
  StepPlan& p = plan(h);
  p.symm.world = (int)recv_ptrs.size();
  for (int i = 0; i < p.symm.world; ++i) p.symm.recv[i] = reinterpret_cast<float*>(recv_ptrs[i]);
  p.symm.step = reinterpret_cast<const uint32_t*>(step);
  p.symm.rank = rank;
  p.symm.n_max = n_max;
}

int g2_num_sms_public();

void dstep_add_gemv(long long h, const void* w, int N, int K, const void* x, int ldx, const void* bias, const void* norm_w, float eps,
                    float norm_offset, int act, const void* residual, void* y, int ldy, bool allreduce) {
  StepPlan& p = plan(h);
  if (p.finalized) throw std::runtime_error("decode_step: plan already finalized");
  if (!gemv2_supported(p.T, K)) throw std::runtime_error("decode_step: activations too wide for shared memory");
  const bool glu = act != 0;
  StepPhase ph{};
  ph.type = PH_GEMV;
  G2Phase& g = ph.g;
  g.x = x; g.bias = bias; g.norm_w = norm_w; g.residual = residual; g.y = y;
  g.T = p.T; g.N = N; g.K = K; g.ldx = ldx; g.ldy = ldy; g.eps = eps; g.norm_offset = norm_offset; g.act = act;
  int rows8, whole, grid, pmax;
  gemv2_plan(N, K, glu, &rows8, &whole, &grid, &pmax);
  g.rows8 = rows8; g.whole_tiles = whole; g.grid = grid; g.p_max = pmax;
  g.parity = -1;
  g.call = 0;
  if (allreduce) {
    if (glu || p.symm.world < 2) throw std::runtime_error("decode_step: all-reduce phase needs a symmetric workspace and a plain epilogue");
    if (N > p.symm.n_max) throw std::runtime_error("decode_step: all-reduce output wider than the workspace");
    g.parity = p.n_ar & 1;
    g.call = p.n_ar;
    p.n_ar += 1;
  }
  ph.expected = grid;
  CUtensorMap tm;
  make_weight_tmap(&tm, w, N, K, (glu || rows8) ? 8 : 16);
  p.tmaps.push_back(tm);
  p.tmap_index.push_back((int)p.tmaps.size() - 1);
  const int n_tiles = gemv2_ntiles(N, glu);
  p.ws_floats = std::max(p.ws_floats, (size_t)n_tiles * pmax * 128);
  p.n_tickets = std::max(p.n_tickets, (size_t)n_tiles);
  const int Kp = (K + G2_KC - 1) / G2_KC * G2_KC;
  p.xs_bytes = std::max(p.xs_bytes, p.T * (Kp * 2 + 64));
  p.phases.push_back(ph);
}

void dstep_add_attn(long long h, const void* qkv, void* out, void* k_cache, void* v_cache, const void* q_norm, const void* k_norm,
                    float norm_eps, int B, int T, int Hq, int Hkv, int D, int S, int L, float scale, int window, const float* sinks,
                    int nsplit) {
  StepPlan& p = plan(h);
  if (p.finalized) throw std::runtime_error("decode_step: plan already finalized");
  if (D != 64 && D != 128) throw std::runtime_error("decode_step: head_dim must be 64 or 128");
  if (p.D != 0 && p.D != D) throw std::runtime_error("decode_step: mixed head dims");
  p.D = D;
  StepPhase ph{};
  ph.type = PH_ATTN;
  AttnArgs& a = ph.a;
  a.k = reinterpret_cast<const __nv_bfloat16*>(k_cache);
  a.v = reinterpret_cast<const __nv_bfloat16*>(v_cache);
  a.k_w = reinterpret_cast<__nv_bfloat16*>(k_cache);
  a.v_w = reinterpret_cast<__nv_bfloat16*>(v_cache);
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a.q_norm = reinterpret_cast<const __nv_bfloat16*>(q_norm);
  a.k_norm = reinterpret_cast<const __nv_bfloat16*>(k_norm);
  a.sinks = sinks;
  a.B = B; a.T = T; a.Hq = Hq; a.Hkv = Hkv; a.S = S; a.L = L; a.nsplit = nsplit; a.window = window;
  a.scale_log2 = scale * kLog2e;
  a.norm_eps = norm_eps;
  a.causal = 1;
  ph.n_items = B * Hkv * nsplit;
  ph.expected = ph.n_items;
  p.max_items = std::max(p.max_items, ph.n_items);
  p.attn_ws_floats = std::max(p.attn_ws_floats, (size_t)B * Hkv * nsplit * 64 * (D + 2));
  p.attn_tickets = std::max(p.attn_tickets, (size_t)B * Hkv);
  p.tmap_index.push_back(-1);
  p.phases.push_back(ph);
}

static void dfree(void* p) {
  if (p) cudaFree(p);
}

void dstep_finalize(long long h) {
  StepPlan& p = plan(h);
  if (p.finalized) return;
  const int sms = g2_num_sms_public();
  if (p.max_items > sms) throw std::runtime_error("decode_step: more attention work items than SMs");
  // shared memory: ring | xs | red | rstd | barriers
  const size_t fixed = (size_t)p.xs_bytes + (G2_CONSUMER_WARPS * 128 + 64) * sizeof(float) + (2 * G2_MAX_STAGES + 1) * sizeof(uint64_t) + 128 + 1024;
  int ns = (int)((G2_SMEM_BUDGET - fixed) / G2_STAGE_BYTES);
  ns = std::min(ns, G2_MAX_STAGES);
  const size_t attn_bytes = p.D ? (size_t)(64 * p.D + 2 * 2 * 64 * p.D) * 2 + (size_t)2 * p.T * p.D * 2 : 0;
  p.attn_stages = (int)((attn_bytes + G2_STAGE_BYTES - 1) / G2_STAGE_BYTES);
  if (ns - p.attn_stages < 2) throw std::runtime_error("decode_step: shared memory too small for ring + attention tiles");
  p.n_stages = ns;
  p.smem = fixed + (size_t)ns * G2_STAGE_BYTES;
  auto ck = [](cudaError_t e) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("decode_step: ") + cudaGetErrorString(e));
  };
  ck(cudaMalloc(&p.d_tmaps, p.tmaps.size() * sizeof(CUtensorMap)));
  ck(cudaMemcpy(p.d_tmaps, p.tmaps.data(), p.tmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  ck(cudaMalloc(&p.d_ws, std::max<size_t>(p.ws_floats, 1) * sizeof(float)));
  ck(cudaMalloc(&p.d_tickets, std::max<size_t>(p.n_tickets, 1) * sizeof(unsigned)));
  ck(cudaMemset(p.d_tickets, 0, std::max<size_t>(p.n_tickets, 1) * sizeof(unsigned)));
  ck(cudaMalloc(&p.d_attn_ws, std::max<size_t>(p.attn_ws_floats, 1) * sizeof(float)));
  ck(cudaMalloc(&p.d_attn_tickets, std::max<size_t>(p.attn_tickets, 1) * sizeof(unsigned)));
  ck(cudaMemset(p.d_attn_tickets, 0, std::max<size_t>(p.attn_tickets, 1) * sizeof(unsigned)));
  for (size_t i = 0; i < p.phases.size(); ++i) {
    StepPhase& ph = p.phases[i];
    if (ph.type == PH_GEMV) {
      ph.g.tmap = reinterpret_cast<const CUtensorMap*>(p.d_tmaps) + p.tmap_index[i];
      ph.g.ws_part = p.d_ws;
      ph.g.tickets = p.d_tickets;
    } else {
      ph.a.ws_o = p.d_attn_ws;
      ph.a.ws_ml = p.d_attn_ws + (size_t)ph.a.B * ph.a.Hkv * ph.a.nsplit * 64 * p.D;
      ph.a.tickets = p.d_attn_tickets;
    }
  }
  ck(cudaMalloc(&p.d_phases, p.phases.size() * sizeof(StepPhase)));
  ck(cudaMemcpy(p.d_phases, p.phases.data(), p.phases.size() * sizeof(StepPhase), cudaMemcpyHostToDevice));
  ck(cudaMalloc(&p.d_done, p.phases.size() * sizeof(unsigned)));
  ck(cudaDeviceSynchronize());
  p.finalized = true;
}

int dstep_num_allreduce(long long h) { return plan(h).n_ar; }

void dstep_free(long long h) {
  StepPlan& p = plan(h);
  dfree(p.d_phases); dfree(p.d_tmaps); dfree(p.d_done); dfree(p.d_ws); dfree(p.d_tickets); dfree(p.d_attn_ws); dfree(p.d_attn_tickets);
  delete plans()[h];
  plans()[h] = nullptr;
}

void dstep_launch(long long h, const int* positions, const int* write_pos, const int* lines, const float* cos, const float* sin,
                  int call_base, int parity_base, cudaStream_t stream) {
  StepPlan& p = plan(h);
  if (!p.finalized) throw std::runtime_error("decode_step: plan not finalized");
  StepParams sp{};
  sp.phases = reinterpret_cast<const StepPhase*>(p.d_phases);
  sp.done = p.d_done;
  sp.n_phases = (int)p.phases.size();
  sp.n_stages = p.n_stages;
  sp.attn_stages = p.attn_stages;
  sp.max_items = p.max_items;
  sp.xs_bytes = p.xs_bytes;
  sp.symm = p.symm;
  sp.max_inflight = (g2_max_inflight() > 0 && g2_max_inflight() < p.n_stages - p.attn_stages) ? g2_max_inflight() : 0;
  sp.call_base = call_base;
  sp.parity_base = parity_base & 1;
  sp.positions = positions; sp.write_pos = write_pos; sp.lines = lines; sp.cos = cos; sp.sin = sin;
  {
    // debug timeline: one slot per phase, contiguous (slots are handed out sequentially)
    long long first = -1;
    for (size_t i = 0; i < p.phases.size(); ++i) {
      const long long sl = prof_next_slot();
      if (i == 0) first = sl;
      if (sl < 0) first = -1;
    }
    sp.prof = prof_slot_ptr(first);
  }
  cudaError_t e = cudaMemsetAsync(p.d_done, 0, p.phases.size() * sizeof(unsigned), stream);
  if (e != cudaSuccess) throw std::runtime_error(std::string("decode_step memset: ") + cudaGetErrorString(e));
  const int sms = g2_num_sms_public();
  if (p.D == 64) {
    auto kern = decode_step_kernel<64>;
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BUDGET); cfg = true; }
    launch_pdl(kern, dim3(sms), dim3(G2_THREADS), p.smem, stream, sp);
  } else {
    auto kern = decode_step_kernel<128>;
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BUDGET); cfg = true; }
    launch_pdl(kern, dim3(sms), dim3(G2_THREADS), p.smem, stream, sp);
  }
}

}  // namespace nxdi
