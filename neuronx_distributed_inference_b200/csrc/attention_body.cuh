// Flash attention for the KV-cache engine: one warp-MMA kernel template serving
//   * decode / speculation over the contiguous cache  [L, Hkv, S, D]   (split-KV, in-kernel combine)
//   * decode over the paged cache                     [nblk, bs, Hkv, D] + block table
//   * causal prefill over fresh K/V                   [B, T, Hkv, D]
// reference kernels: K1 attention_cte (prefill; causal / sliding window / sink), K2 attention_block_tkg
// (decode attention stage), K13 sliding-window flash_fwd (modules/sliding_window/attention.py:235-477).
//
// A CTA = 4 warps works on up to 64 "rows" that share one KV head:
//   decode : rows = (active token t, q head g of the GQA group)   -> K/V read ONCE per group
//   prefill: rows = 64 consecutive tokens of one q head
// Row blocks of 16 map to warps (RBp in {1,2,4}); the remaining factor KS = 4/RBp splits the 64 keys of each
// K/V tile between warps, so a 4-row decode step still uses all four warps.  K/V tiles stream through a
// 4-stage cp.async pipeline into XOR-swizzled shared memory; QK^T and PV run on mma.sync m16n8k16 (bf16, fp32
// accumulate) with ldmatrix operand loads; softmax is online in the exp2 domain.  Split-KV partials go to a
// workspace and the last CTA of a (batch, kv-head) — elected by an atomic ticket — combines them, so decode
// attention is ONE launch regardless of context length.
#pragma once
#include <cfloat>

#include "api.h"
#include "common.cuh"

namespace nxdi {

enum { ATTN_DECODE = 0, ATTN_PAGED = 1, ATTN_PREFILL = 2 };
constexpr int ATT_TILE = 64;
constexpr int ATT_THREADS = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnArgs {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* out;
  const int* lines;
  const int* positions;
  const int* block_table;
  const float* sinks;
  float* ws_o;
  float* ws_ml;
  unsigned* tickets;
  int B, T, Hq, Hkv, S, L, nsplit, window, block_size, max_blocks;
  float scale_log2;
  // fused decode prologue (ATTN_DECODE only): q/k/v come straight from the QKV projection; per-head RMSNorm + RoPE are
  // applied here, the new K/V rows are written to the cache by this CTA before it reads the cache
  const __nv_bfloat16* qkv;     // [B*T, (Hq + 2 Hkv) * D] or null
  const float* cos;             // [B*T, D/2]
  const float* sin;
  const __nv_bfloat16* q_norm;  // [D] or null
  const __nv_bfloat16* k_norm;
  const int* write_pos;         // [B*T] cache slot of every active token (-1 = skip)
  __nv_bfloat16* k_w;
  __nv_bfloat16* v_w;
  float norm_eps;
  unsigned long long* prof;  // debug timeline or null
  int causal;  // prefill: 1 = causal (default), 0 = bidirectional (encoders: vision towers, Whisper, diffusion)
};

template <int D>
__device__ __forceinline__ int swz(int row, int chunk) {  // element offset of a 16-byte chunk in a [rows][D] bf16 tile
  return row * D + ((chunk ^ (row & 7)) << 3);
}

// ATT_STAGES: cp.async ring depth.  3 for short decode contexts (latency bound: 3 x 64 keys in flight; 112 KB at D=128 so that
// the CTA still co-resides with a ~108 KB decode GEMV CTA whose producer is prefetching the next projection's weights);
// 2 elsewhere (80 KB of shared memory -> two CTAs per SM, which the bandwidth/FLOP-bound cases need).
// The body is a device function so that the persistent decode-step kernel (decode_step.cu) runs the very same code as one of
// its phases: `tid` is the thread's index inside the 128-thread group that executes the work item, (bx, by, gdx) the item's
// coordinates in the (batch x kv-head, split) grid, `bar_id` the named barrier of the group (0 == __syncthreads for a
// 128-thread CTA), `standalone` whether the programmatic-dependent-launch trigger belongs to this code; `wait_dep()` blocks until the
// producer of q / qkv has finished (everything before it only touches step inputs and cache rows of earlier steps).
template <int D, int MODE, int ATT_STAGES, class WaitDep>
__device__ __forceinline__ void attention_body(const AttnArgs& p, uint8_t* smem_raw, const int tid, const int bx, const int by,
                                               const int gdx, const int bar_id, const bool standalone, WaitDep&& wait_dep) {
  constexpr int CH = D / 8;  // 16-byte chunks per row
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(ATT_THREADS) : "memory"); };
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // [64][D]
  __nv_bfloat16* sK = sQ + 64 * D;                                 // [STAGES][64][D]
  __nv_bfloat16* sV = sK + ATT_STAGES * 64 * D;                    // [STAGES][64][D]
  __nv_bfloat16* sNew = sV + ATT_STAGES * 64 * D;                  // fused decode: [2 (k,v)][T][D] rows of this step
  __shared__ float sM[4][16], sL[4][16];
  __shared__ int s_pos[64];
  __shared__ int s_wp[64];
  __shared__ bool s_last;

  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int G = p.Hq / p.Hkv;

  // Fused decode (q/k/v straight from the QKV projection) splits the kernel at the dependency: everything that only needs
  // the step's inputs (positions, cache lines) and the cache rows of EARLIER steps — the tile ranges, the first ring fill of
  // K/V tiles, cos/sin — runs BEFORE griddepcontrol.wait and overlaps the QKV projection; after the wait only the new rows are
  // fetched, rotated, written to the cache (for later steps) and patched into the resident tiles from shared memory.
  const bool fused = (MODE == ATTN_DECODE) && p.qkv != nullptr;
  unsigned long long* prof = p.prof ? p.prof + (size_t)(bx + by * gdx) * 8 : nullptr;
  if (prof && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prof[0] = t;
    prof[1] = clock64();
  }
  if (standalone) pdl_launch_dependents();
  if (!fused) wait_dep();

  // ---- work decomposition ----------------------------------------------------------------------------
  int b, kvh, split = 0, R, qh0 = 0, tok0 = 0;
  if (MODE == ATTN_PREFILL) {
    const int n_qt = (p.T + 63) / 64;
    const int qt = n_qt - 1 - (int)(bx % n_qt);  // heavy (late) tiles first
    const int bh = bx / n_qt;
    b = bh / p.Hq;
    qh0 = bh % p.Hq;
    kvh = qh0 / G;
    tok0 = qt * 64;
    R = min(64, p.T - tok0);
  } else {
    b = bx / p.Hkv;
    kvh = bx % p.Hkv;
    split = by;
    R = p.T * G;
  }
  const int RB = (R + 15) >> 4;
  const int RBp = RB <= 1 ? 1 : (RB <= 2 ? 2 : 4);
  const int KS = 4 / RBp;
  const int rb = warp % RBp, ks = warp / RBp;
  const bool warp_active = rb < RB;

  // per-row absolute positions (key j visible iff j <= pos and j > pos - window)
  int pos_max = -1, pos_min = 0x7fffffff;
  for (int r = tid; r < 64; r += ATT_THREADS) {
    int ps = -1;
    if (r < R) ps = (MODE == ATTN_PREFILL) ? (p.causal ? tok0 + r : p.T - 1) : p.positions[b * p.T + r / G];
    s_pos[r] = ps;
  }
  group_sync();
  for (int r = 0; r < R; ++r) {
    pos_max = max(pos_max, s_pos[r]);
    pos_min = min(pos_min, s_pos[r]);
  }
  const int cap = (MODE == ATTN_PREFILL) ? p.T : (MODE == ATTN_PAGED ? p.max_blocks * p.block_size : p.S);
  const int kv_len = min(pos_max + 1, cap);
  int line = 0;
  bool seq_ok = true;
  if (MODE == ATTN_DECODE) {
    line = p.lines[b];
    seq_ok = line >= 0 && line < p.L;
  }
  const int nt = seq_ok ? (kv_len + ATT_TILE - 1) / ATT_TILE : 0;
  int t_lo = 0;
  if (p.window > 0) t_lo = max(0, pos_min - p.window + 1) / ATT_TILE;
  const int span = max(nt - t_lo, 0);
  const int tps = (span + p.nsplit - 1) / p.nsplit;
  const int t_beg = t_lo + split * tps;
  const int t_end = min(nt, t_beg + tps);

  auto load_tile = [&](int tile, int buf) {
    __nv_bfloat16* dK = sK + buf * 64 * D;
    __nv_bfloat16* dV = sV + buf * 64 * D;
    for (int i = tid; i < 64 * CH; i += ATT_THREADS) {
      const int r = i / CH, c = i % CH;
      const int key = tile * ATT_TILE + r;
      const bool ok = key < kv_len;
      const int kk = ok ? key : 0;
      size_t off;
      if (MODE == ATTN_DECODE) {
        off = (((size_t)line * p.Hkv + kvh) * p.S + kk) * D;
      } else if (MODE == ATTN_PAGED) {
        const int blk = p.block_table[(size_t)b * p.max_blocks + kk / p.block_size];
        off = (((size_t)max(blk, 0) * p.block_size + kk % p.block_size) * p.Hkv + kvh) * D;
      } else {
        off = (((size_t)b * p.T + kk) * p.Hkv + kvh) * D;
      }
      cp_async16(dK + swz<D>(r, c), p.k + off + c * 8, ok);
      cp_async16(dV + swz<D>(r, c), p.v + off + c * 8, ok);
    }
  };
  auto prefetch_ring = [&]() {
#pragma unroll
    for (int s = 0; s < ATT_STAGES - 1; ++s) {
      if (t_beg + s < t_end) load_tile(t_beg + s, s);
      cp_async_commit();  // one group per tile slot (possibly empty) keeps the wait arithmetic uniform
    }
  };

  // ---- fused prologue: RMSNorm + RoPE on q (-> sQ) and on the new k rows; k/v appended to the cache ----------
  // Every split-CTA of this (batch, kv head) computes the same new rows (identical values) and patches its own tiles, so no
  // inter-CTA ordering is needed; the global write only serves later steps.
  int new_tile_lo = 0x7fffffff, new_tile_hi = -1;
  if (fused) {
    constexpr int HALF = D / 2, PPL = HALF / 32;
    prefetch_ring();                                         // old rows: safe before the dependency
    for (int t = tid; t < p.T; t += ATT_THREADS) {
      const int wp = p.write_pos[(size_t)b * p.T + t];
      s_wp[t] = (seq_ok && wp >= 0 && wp < p.S) ? wp : -1;
    }
    const int nrows = R + 2 * p.T;
    for (int i = tid; i < (64 - R) * CH; i += ATT_THREADS)   // rows past R stay zero
      *reinterpret_cast<uint4*>(sQ + swz<D>(R + i / CH, i % CH)) = make_uint4(0u, 0u, 0u, 0u);
    bool waited = false;
    for (int base = 0; base < nrows; base += 16) {           // 4 warps x 4 rows per pass, loads of a pass issued together
      float x1[4][PPL], x2[4][PPL], cs[4][PPL], sn[4][PPL];
      int t_[4], head_[4], kind_[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = base + warp + 4 * u;
        int t = 0, head = 0, kind = 3;
        if (row < R) { t = row / G; head = kvh * G + row % G; kind = 0; }
        else if (row < R + p.T) { t = row - R; head = p.Hq + kvh; kind = 1; }
        else if (row < nrows) { t = row - R - p.T; head = p.Hq + p.Hkv + kvh; kind = 2; }
        t_[u] = t; head_[u] = head; kind_[u] = kind;
        if (kind < 2) {
          const size_t bt = (size_t)b * p.T + t;
#pragma unroll
          for (int j = 0; j < PPL; ++j) {
            cs[u][j] = p.cos[bt * HALF + lane + 32 * j];
            sn[u][j] = p.sin[bt * HALF + lane + 32 * j];
          }
        }
      }
      if (!waited) { wait_dep(); waited = true; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kind_[u] < 3) {
          const __nv_bfloat16* src = p.qkv + (((size_t)b * p.T + t_[u]) * (p.Hq + 2 * p.Hkv) + head_[u]) * D;
#pragma unroll
          for (int j = 0; j < PPL; ++j) {
            x1[u][j] = ldg_act_bf16(src + lane + 32 * j);
            x2[u][j] = ldg_act_bf16(src + lane + 32 * j + HALF);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kind = kind_[u], t = t_[u];
        if (kind == 3) continue;
        const int row = base + warp + 4 * u;
        if (kind != 2) {
          const __nv_bfloat16* nw = kind == 0 ? p.q_norm : p.k_norm;
          if (nw != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < PPL; ++j) ss += x1[u][j] * x1[u][j] + x2[u][j] * x2[u][j];
            ss = warp_sum(ss);
            const float rstd = rsqrtf(ss / (float)D + p.norm_eps);
#pragma unroll
            for (int j = 0; j < PPL; ++j) {   // rounded to bf16 before the rotation, like the unfused path
              x1[u][j] = __bfloat162float(__float2bfloat16(x1[u][j] * rstd * __bfloat162float(nw[lane + 32 * j])));
              x2[u][j] = __bfloat162float(__float2bfloat16(x2[u][j] * rstd * __bfloat162float(nw[lane + 32 * j + HALF])));
            }
          }
#pragma unroll
          for (int j = 0; j < PPL; ++j) {
            const float o1 = x1[u][j] * cs[u][j] - x2[u][j] * sn[u][j], o2 = x2[u][j] * cs[u][j] + x1[u][j] * sn[u][j];
            x1[u][j] = o1;
            x2[u][j] = o2;
          }
        }
        if (kind == 0) {
#pragma unroll
          for (int j = 0; j < PPL; ++j) {
            const int c1 = lane + 32 * j, c2 = c1 + HALF;
            sQ[swz<D>(row, c1 >> 3) + (c1 & 7)] = __float2bfloat16(x1[u][j]);
            sQ[swz<D>(row, c2 >> 3) + (c2 & 7)] = __float2bfloat16(x2[u][j]);
          }
        } else {
          __nv_bfloat16* sn_row = sNew + ((size_t)(kind - 1) * p.T + t) * D;
          const int wp = p.write_pos[(size_t)b * p.T + t];
          const bool wr = seq_ok && wp >= 0 && wp < p.S;
          __nv_bfloat16* dst = (kind == 1 ? p.k_w : p.v_w) + (((size_t)line * p.Hkv + kvh) * p.S + (wr ? wp : 0)) * D;
#pragma unroll
          for (int j = 0; j < PPL; ++j) {
            const __nv_bfloat16 y1 = __float2bfloat16(x1[u][j]), y2 = __float2bfloat16(x2[u][j]);
            sn_row[lane + 32 * j] = y1;
            sn_row[lane + 32 * j + HALF] = y2;
            if (wr) {
              dst[lane + 32 * j] = y1;
              dst[lane + 32 * j + HALF] = y2;
            }
          }
        }
      }
    }
    if (!waited) wait_dep();
    if (prof && tid == 0) prof[2] = clock64();   // (last row pass of this thread done, not the wait itself)
    group_sync();   // sQ, sNew, s_wp visible
    if (prof && tid == 0) prof[3] = clock64();
    for (int t = 0; t < p.T; ++t) {
      const int wp = s_wp[t];
      if (wp >= 0) {
        new_tile_lo = min(new_tile_lo, wp / ATT_TILE);
        new_tile_hi = max(new_tile_hi, wp / ATT_TILE);
      }
    }
  }

  // ---- Q tile -> shared (swizzled), then A fragments ----------------------------------------------------
  for (int i = tid; i < (fused ? 0 : 64 * CH); i += ATT_THREADS) {
    const int r = i / CH, c = i % CH;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < R) {
      const __nv_bfloat16* src;
      if (MODE == ATTN_PREFILL)
        src = p.q + (((size_t)b * p.T + tok0 + r) * p.Hq + qh0) * D;
      else
        src = p.q + (((size_t)b * p.T + r / G) * p.Hq + kvh * G + r % G) * D;
      val = ldg_act(src + c * 8);
    }
    *reinterpret_cast<uint4*>(sQ + swz<D>(r, c)) = val;
  }

  if (!fused) prefetch_ring();
  group_sync();  // sQ visible

  uint32_t qf[D / 16][4];
  if (warp_active) {
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      const int row = rb * 16 + (lane & 15);
      const int chunk = kk * 2 + (lane >> 4);
      ldmatrix_x4(qf[kk], sQ + swz<D>(row, chunk));
    }
  }
  const int pos_a = s_pos[min(rb * 16 + g, 63)], pos_b = s_pos[min(rb * 16 + g + 8, 63)];
  int rb_pos_max = -1;  // newest position any row of my block can see (causal skip of whole sub-tiles)
  for (int i = 0; i < 16; ++i) rb_pos_max = max(rb_pos_max, s_pos[min(rb * 16 + i, 63)]);

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

  for (int tile = t_beg; tile < t_end; ++tile) {
    const int buf = (tile - t_beg) % ATT_STAGES;
    if (tile + ATT_STAGES - 1 < t_end) load_tile(tile + ATT_STAGES - 1, (tile - t_beg + ATT_STAGES - 1) % ATT_STAGES);
    cp_async_commit();
    cp_async_wait<ATT_STAGES - 1>();
    group_sync();
    if (fused && tile >= new_tile_lo && tile <= new_tile_hi) {
      // this step's K/V rows: the cp.async above may have fetched them stale (or not at all) — take them from shared memory
      __nv_bfloat16* pK = sK + buf * 64 * D;
      __nv_bfloat16* pV = sV + buf * 64 * D;
      for (int i = tid; i < p.T * 2 * CH; i += ATT_THREADS) {
        const int t = i / (2 * CH), rem = i % (2 * CH), which = rem / CH, c = rem % CH;
        const int r = s_wp[t] - tile * ATT_TILE;
        if (s_wp[t] >= 0 && r >= 0 && r < ATT_TILE)
          *reinterpret_cast<uint4*>((which ? pV : pK) + swz<D>(r, c)) =
              *reinterpret_cast<const uint4*>(sNew + ((size_t)which * p.T + t) * D + c * 8);
      }
      group_sync();
    }
    if (warp_active) {
      const __nv_bfloat16* tK = sK + buf * 64 * D;
      const __nv_bfloat16* tV = sV + buf * 64 * D;
      for (int j = 0; j < RBp; ++j) {
        const int st = ks * RBp + j;  // 16-key sub-tile inside the 64-key tile
        const int key0 = tile * ATT_TILE + st * 16;
        if (key0 >= kv_len || key0 > rb_pos_max) break;  // nothing visible to this row block from here on
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t kf[4];
          const int row = st * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int chunk = kk * 2 + ((lane >> 3) & 1);
          ldmatrix_x4(kf, tK + swz<D>(row, chunk));
          const uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
          mma_bf16_16816(s0, qf[kk], b0);
          mma_bf16_16816(s1, qf[kk], b1);
        }
        // s0: keys key0 + 2*t4 + {0,1}; s1: keys key0 + 8 + 2*t4 + {0,1}; regs [0,1] row g, [2,3] row g+8
        float sc[8] = {s0[0], s0[1], s1[0], s1[1], s0[2], s0[3], s1[2], s1[3]};
        const int kidx[4] = {key0 + 2 * t4, key0 + 2 * t4 + 1, key0 + 8 + 2 * t4, key0 + 9 + 2 * t4};
        float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kj = kidx[e];
          const bool va = kj <= pos_a && kj < kv_len && (p.window <= 0 || kj > pos_a - p.window);
          const bool vb = kj <= pos_b && kj < kv_len && (p.window <= 0 || kj > pos_b - p.window);
          sc[e] = va ? sc[e] * p.scale_log2 : -INFINITY;
          sc[4 + e] = vb ? sc[4 + e] * p.scale_log2 : -INFINITY;
          mx_a = fmaxf(mx_a, sc[e]);
          mx_b = fmaxf(mx_b, sc[4 + e]);
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
        const float ca = (mn_a == -INFINITY) ? 1.f : exp2f(m_a - mn_a);
        const float cb = (mn_b == -INFINITY) ? 1.f : exp2f(m_b - mn_b);
        const float ba = (mn_a == -INFINITY) ? 0.f : mn_a, bb = (mn_b == -INFINITY) ? 0.f : mn_b;
        float pa[4], pb[4], sa = 0.f, sb = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pa[e] = exp2f(sc[e] - ba);
          pb[e] = exp2f(sc[4 + e] - bb);
          sa += pa[e];
          sb += pb[e];
        }
        l_a = l_a * ca + sa;
        l_b = l_b * cb + sb;
        m_a = mn_a;
        m_b = mn_b;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
          o[i][0] *= ca;
          o[i][1] *= ca;
          o[i][2] *= cb;
          o[i][3] *= cb;
        }
        const uint32_t pf[4] = {pack_bf16(pa[0], pa[1]), pack_bf16(pb[0], pb[1]), pack_bf16(pa[2], pa[3]),
                                pack_bf16(pb[2], pb[3])};
#pragma unroll
        for (int dd = 0; dd < D / 16; ++dd) {
          uint32_t vf[4];
          const int row = st * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
          const int chunk = dd * 2 + (lane >> 4);
          ldmatrix_x4_trans(vf, tV + swz<D>(row, chunk));
          const uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
          mma_bf16_16816(o[dd * 2], pf, b0);
          mma_bf16_16816(o[dd * 2 + 1], pf, b1);
        }
      }
    }
    group_sync();
  }

  if (prof && tid == 0) prof[4] = clock64();
  // ---- merge the KS key-split warps of each row block (through shared memory, reusing the K buffers) ----
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  float* sO = reinterpret_cast<float*>(sK);  // [4 warps][16][D] fp32 = 4*16*D*4 bytes <= size of sK+sV
  if (t4 == 0) {
    sM[warp][g] = m_a;
    sM[warp][g + 8] = m_b;
    sL[warp][g] = l_a;
    sL[warp][g + 8] = l_b;
  }
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    float* base = sO + (size_t)warp * 16 * D;
    *reinterpret_cast<float2*>(base + g * D + i * 8 + 2 * t4) = make_float2(o[i][0], o[i][1]);
    *reinterpret_cast<float2*>(base + (g + 8) * D + i * 8 + 2 * t4) = make_float2(o[i][2], o[i][3]);
  }
  group_sync();

  const int bh = (MODE == ATTN_PREFILL) ? 0 : bx;
  const bool direct = (MODE == ATTN_PREFILL) || p.nsplit == 1;
  // per-row merge factors once (row max over the KS key-split warps, rescale factor of each, denominator incl. the sink):
  // sF[k2][r] = exp2(m_k2 - M) (/ L when written directly), so the element loop below is KS multiply-adds per output
  float* sF = reinterpret_cast<float*>(sQ);   // [4][64] — the Q tile is dead after the fragments were loaded
  float* sMl = sF + 4 * 64;                   // [64][2]: row max, row denominator (split-KV partials)
  for (int r = tid; r < RB * 16; r += ATT_THREADS) {
    const int rbi = r >> 4, i = r & 15;
    float M = -INFINITY;
    for (int k2 = 0; k2 < KS; ++k2) M = fmaxf(M, sM[rbi + RBp * k2][i]);
    float Lsum = 0.f, f[4] = {0.f, 0.f, 0.f, 0.f};
    if (M != -INFINITY) {
      for (int k2 = 0; k2 < KS; ++k2) {
        f[k2] = exp2f(sM[rbi + RBp * k2][i] - M);
        Lsum += sL[rbi + RBp * k2][i] * f[k2];
      }
    }
    float inv = 1.f;
    if (direct) {
      float Lf = Lsum;
      if (p.sinks != nullptr && r < R) {
        const int qh = (MODE == ATTN_PREFILL) ? qh0 : kvh * G + r % G;
        const float sk = p.sinks[qh] * kLog2e;
        if (M == -INFINITY) Lf = 1.f; else Lf += exp2f(sk - M);
      }
      inv = (M == -INFINITY || Lf == 0.f) ? 0.f : 1.f / Lf;
    }
    for (int k2 = 0; k2 < 4; ++k2) sF[k2 * 64 + r] = f[k2] * inv;
    sMl[r * 2] = M;
    sMl[r * 2 + 1] = Lsum;
  }
  group_sync();
  for (int e = tid; e < R * D; e += ATT_THREADS) {
    const int r = e / D, d = e % D;
    const int rbi = r >> 4, i = r & 15;
    float acc = 0.f;
    for (int k2 = 0; k2 < KS; ++k2) acc += sO[((size_t)(rbi + RBp * k2) * 16 + i) * D + d] * sF[k2 * 64 + r];
    if (direct) {
      int qh, tok;
      if (MODE == ATTN_PREFILL) {
        qh = qh0;
        tok = tok0 + r;
      } else {
        qh = kvh * G + r % G;
        tok = r / G;
      }
      p.out[(((size_t)b * p.T + tok) * p.Hq + qh) * D + d] = __float2bfloat16(acc);
    } else {
      const size_t wrow = ((size_t)bh * p.nsplit + split) * 64 + r;
      p.ws_o[wrow * D + d] = acc;
      if (d == 0) {
        p.ws_ml[wrow * 2] = sMl[r * 2];
        p.ws_ml[wrow * 2 + 1] = sMl[r * 2 + 1];
      }
    }
  }
  if (prof && tid == 0) {
    unsigned long long t;
    unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    prof[5] = clock64();
    prof[6] = t;
    prof[7] = sm;
  }
  if (direct) return;

  // ---- split-KV combine by the last-arriving CTA of this (batch, kv head) ------------------------------------
  __threadfence();
  group_sync();
  if (tid == 0) s_last = (atomicAdd(&p.tickets[bh], 1u) == (unsigned)(p.nsplit - 1));
  group_sync();
  if (!s_last) return;
  __threadfence();
  for (int e = tid; e < R * D; e += ATT_THREADS) {
    const int r = e / D, d = e % D;
    float M = -INFINITY;
    for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(p.ws_ml + (((size_t)bh * p.nsplit + s) * 64 + r) * 2));
    float Lsum = 0.f, acc = 0.f;
    if (M != -INFINITY) {
      for (int s = 0; s < p.nsplit; ++s) {
        const size_t wrow = ((size_t)bh * p.nsplit + s) * 64 + r;
        const float ms = __ldcg(p.ws_ml + wrow * 2);
        if (ms == -INFINITY) continue;
        const float f = exp2f(ms - M);
        Lsum += __ldcg(p.ws_ml + wrow * 2 + 1) * f;
        acc += __ldcg(p.ws_o + wrow * D + d) * f;
      }
    }
    const int qh = kvh * G + r % G, tok = r / G;
    if (p.sinks != nullptr) {
      const float sk = p.sinks[qh] * kLog2e;
      if (M == -INFINITY) Lsum = 1.f; else Lsum += exp2f(sk - M);
    }
    const float val = (M == -INFINITY || Lsum == 0.f) ? 0.f : acc / Lsum;
    p.out[(((size_t)b * p.T + tok) * p.Hq + qh) * D + d] = __float2bfloat16(val);
  }
  if (tid == 0) p.tickets[bh] = 0;  // re-arm for the next launch / graph replay
}

}  // namespace nxdi
