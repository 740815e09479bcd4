// pybind11 surface of the sm_100a extension (neuronx_distributed_inference_b200._C).  Compiled by the host
// compiler; the CUDA translation units expose raw-pointer launchers (api.h).
#include <limits>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <map>

#include "api.h"

namespace nxdi {

#define NXDI_CUDA_OK(call)                                                        \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    TORCH_CHECK(e__ == cudaSuccess, #call, " failed: ", cudaGetErrorString(e__)); \
  } while (0)

static inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }
static inline const void* optr(const c10::optional<at::Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }
static inline bool is_bf16(const at::Tensor& t) { return t.scalar_type() == at::kBFloat16; }

// ---- rmsnorm ------------------------------------------------------------------------------------------
std::tuple<at::Tensor, at::Tensor> rmsnorm(const at::Tensor& x, const at::Tensor& w, double eps, double offset,
                                           const c10::optional<at::Tensor>& residual) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && is_bf16(x) && x.is_contiguous() && is_bf16(w));
  const int rows = x.size(0), H = x.size(1);
  TORCH_CHECK(H % 8 == 0 && H <= 16384, "rmsnorm: hidden must be a multiple of 8 and <= 16384");
  c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty_like(x);
  at::Tensor res_out;
  if (residual.has_value()) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == x.sizes() && is_bf16(*residual));
    res_out = at::empty_like(x);
  }
  rmsnorm_launch(x.data_ptr(), optr(residual), w.data_ptr(), y.data_ptr(), res_out.defined() ? res_out.data_ptr() : nullptr,
                 rows, H, (float)eps, (float)offset, cur_stream());
  return {y, res_out.defined() ? res_out : y};
}

struct Scratch {
  at::Tensor f, i, tickets, gemv_ws, gemv_tickets;
  // Buffers that were outgrown are RETIRED, never freed: CUDA graphs captured earlier hold their addresses (a decode graph of a
  // small sequence bucket keeps replaying after a larger bucket made the workspace grow); freeing them would hand the memory to
  // the caching allocator while those graphs still write their split-KV partials / stream-K tickets there.
  std::vector<at::Tensor> retired;
  void grow(at::Tensor& t, int64_t n, const at::TensorOptions& o, bool zero) {
    if (t.defined() && t.numel() >= n) return;
    if (t.defined()) retired.push_back(t);
    t = zero ? at::zeros({n}, o) : at::empty({n}, o);
  }
};
static Scratch& scratch(const at::Device& dev, int64_t nf, int64_t ni, int64_t nt);

static void run_gemv(GemvParams& p, int mode, const at::Device& dev) {
  const int wt = p.scale != nullptr ? p.wdtype : 0;
  if (gemv2_supported(p.T, p.K, wt) && p.x_in_smem) {
    auto& s = scratch(dev, 0, 0, 0);
    const bool glu = p.act != 0;
    const int n_tiles = gemv2_ntiles(p.N, glu);
    const int64_t need = (int64_t)n_tiles * gemv2_pmax(p.N, p.K, glu, wt) * 128;
    auto o = at::TensorOptions().device(dev);
    s.grow(s.gemv_ws, std::max<int64_t>(need, 4 << 20), o.dtype(at::kFloat), false);
    s.grow(s.gemv_tickets, std::max<int64_t>(n_tiles, 1 << 15), o.dtype(at::kInt), true);
    gemv2_launch(p, mode, s.gemv_ws.data_ptr<float>(), reinterpret_cast<unsigned*>(s.gemv_tickets.data_ptr<int>()), cur_stream());
  } else {
    TORCH_CHECK(wt == 0, "gemv: 8-bit weights need the TMA streaming kernel");
    TORCH_CHECK(p.K % 256 == 0, "gemv: K must be a multiple of 256 on the wide-activation fallback");
    gemv_launch(p, mode, cur_stream());
  }
}

// ---- skinny GEMM ----------------------------------------------------------------------------------------
static void fill_params(GemvParams& p, const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& bias,
                        const c10::optional<at::Tensor>& norm_w, double eps, double offset, int act,
                        const c10::optional<at::Tensor>& residual, const c10::optional<at::Tensor>& scale, at::Tensor& y,
                        bool x_in_smem) {
  p.x = x.data_ptr();
  p.w = w.data_ptr();
  p.bias = optr(bias);
  p.norm_w = x_in_smem ? optr(norm_w) : nullptr;
  p.residual = optr(residual);
  p.scale = scale.has_value() ? scale->data_ptr<float>() : nullptr;
  p.scale_n = scale.has_value() ? (int)scale->numel() : 0;
  p.y = y.data_ptr();
  p.T = x.size(0);
  p.K = x.size(1);
  p.N = w.size(0);
  p.ldx = x.stride(0);
  p.ldy = y.stride(0);
  p.eps = (float)eps;
  p.norm_offset = (float)offset;
  p.act = act;
  p.x_in_smem = x_in_smem ? 1 : 0;
  p.wdtype = w.scalar_type() == at::kBFloat16 ? 0 : (w.scalar_type() == at::kChar ? 1 : 2);
}

static at::Tensor prenorm_if_needed(const at::Tensor& x, const c10::optional<at::Tensor>& norm_w, double eps, double offset,
                                    bool& x_in_smem) {
  x_in_smem = gemv_smem_bytes(x.size(0), x.size(1), true) <= 220 * 1024;
  if (!x_in_smem && norm_w.has_value()) return std::get<0>(rmsnorm(x.contiguous(), *norm_w, eps, offset, c10::nullopt));
  return x;
}

static void check_gemv_inputs(const at::Tensor& x, const at::Tensor& w) {
  TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1));
  TORCH_CHECK(is_bf16(x) && is_bf16(w), "gemv: bf16 activations and weights");
  TORCH_CHECK(x.size(0) >= 1 && x.size(0) <= GEMV_MAX_T, "gemv: 1..8 tokens");
  TORCH_CHECK(x.size(1) % 64 == 0 && x.stride(1) == 1 && x.stride(0) % 8 == 0 && w.is_contiguous());
}

at::Tensor gemv(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& bias,
                const c10::optional<at::Tensor>& norm_w, double eps, double offset, int64_t act,
                const c10::optional<at::Tensor>& scale, const c10::optional<at::Tensor>& residual) {
  if (scale.has_value()) {
    // weight-only quantised path (int8 / fp8-e4m3, per-channel or per-tensor scale)
    TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && is_bf16(x));
    TORCH_CHECK(w.is_contiguous() && x.stride(1) == 1 && x.size(0) >= 1 && x.size(0) <= GEMV_MAX_T && x.size(1) % 16 == 0);
    const bool i8 = w.scalar_type() == at::kChar, f8 = w.scalar_type() == at::kFloat8_e4m3fn;
    TORCH_CHECK(i8 || f8, "qgemv: int8 or float8_e4m3fn weights");
    TORCH_CHECK(scale->scalar_type() == at::kFloat && scale->is_contiguous() && (scale->numel() == 1 || scale->numel() == w.size(0)));
    c10::cuda::CUDAGuard guard(x.device());
    const int N = w.size(0);
    const bool glu = act != 0;
    auto y = at::empty({x.size(0), glu ? N / 2 : N}, x.options());
    if (residual.has_value()) TORCH_CHECK(!glu && residual->is_contiguous() && residual->size(1) == N && is_bf16(*residual));
    // the TMA-streamed kernel (same ring as bf16, bytes expanded in registers) whenever the shapes allow it
    static const bool q2 = []() { const char* e = getenv("NXDI_B200_QGEMV2"); return e == nullptr || atoi(e) != 0; }();
    if (q2 && gemv2_supported((int)x.size(0), (int)x.size(1), i8 ? 1 : 2) && x.stride(0) % 8 == 0 && (!glu || N % 2 == 0)) {
      GemvParams p{};
      fill_params(p, x, w, bias, norm_w, eps, offset, (int)act, residual, scale, y, true);
      run_gemv(p, 0, x.device());
      return y;
    }
    static int n_sms = 0;
    if (n_sms == 0) n_sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    qgemv_launch(x.data_ptr(), w.data_ptr(), scale->data_ptr<float>(), (int)scale->numel(), optr(bias), optr(norm_w), optr(residual),
                 y.data_ptr(), x.size(0), N, x.size(1), x.stride(0), y.stride(0), (int)act, i8 ? 1 : 2, (float)eps, (float)offset,
                 n_sms, cur_stream());
    return y;
  }
  check_gemv_inputs(x, w);
  c10::cuda::CUDAGuard guard(x.device());
  const int N = w.size(0);
  const bool glu = act != 0;
  if (glu) TORCH_CHECK(N % 2 == 0);
  bool x_in_smem;
  at::Tensor xn = prenorm_if_needed(x, norm_w, eps, offset, x_in_smem);
  auto y = at::empty({x.size(0), glu ? N / 2 : N}, x.options());
  if (residual.has_value())
    TORCH_CHECK(!glu && residual->is_contiguous() && residual->size(0) == x.size(0) && residual->size(1) == N && is_bf16(*residual));
  GemvParams p{};
  fill_params(p, xn, w, bias, norm_w, eps, offset, (int)act, residual, scale, y, x_in_smem);
  run_gemv(p, 0, x.device());
  return y;
}

// tcgen05 GEMM for T > 8 tokens (prefill, long speculation windows)
at::Tensor gemm(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& bias, int64_t act,
                const c10::optional<at::Tensor>& residual, const c10::optional<at::Tensor>& out) {
  TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1));
  TORCH_CHECK(is_bf16(x) && is_bf16(w) && x.stride(1) == 1 && w.is_contiguous());
  const int M = x.size(0), K = x.size(1), N = w.size(0);
  TORCH_CHECK(K % 64 == 0 && x.stride(0) % 8 == 0, "gemm: K must be a multiple of 64");
  const bool glu = act != 0;
  const int n_out = glu ? N / 2 : N;
  TORCH_CHECK(n_out % 8 == 0 && (!glu || N % 2 == 0));
  c10::cuda::CUDAGuard guard(x.device());
  if (out.has_value())
    TORCH_CHECK(out->is_cuda() && is_bf16(*out) && out->dim() == 2 && out->size(0) == M && out->size(1) == n_out && out->is_contiguous());
  auto y = out.has_value() ? *out : at::empty({M, n_out}, x.options());
  if (residual.has_value())
    TORCH_CHECK(!glu && residual->is_contiguous() && residual->size(0) == M && residual->size(1) == n_out && is_bf16(*residual));
  gemm_tcgen05_launch(x.data_ptr(), (int)x.stride(0), w.data_ptr(), optr(bias), optr(residual), y.data_ptr(), n_out, M, N, K,
                      (int)act, cur_stream());
  return y;
}

// Row-parallel GEMV -> one-shot all-reduce over NVLink peer buffers -> +bias +residual.  ONE kernel.
at::Tensor gemv_allreduce(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& bias,
                          const c10::optional<at::Tensor>& residual, const std::vector<int64_t>& recv_ptrs,
                          const at::Tensor& step, int64_t rank, int64_t parity, int64_t call, int64_t n_max,
                          const c10::optional<at::Tensor>& scale) {
  int wt = 0;
  if (scale.has_value()) {   // weight-only int8 / fp8 row-parallel layer
    TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && is_bf16(x) && w.is_contiguous());
    TORCH_CHECK(x.size(0) >= 1 && x.size(0) <= GEMV_MAX_T && x.stride(1) == 1 && x.stride(0) % 8 == 0);
    wt = w.scalar_type() == at::kChar ? 1 : 2;
    TORCH_CHECK(wt == 1 || w.scalar_type() == at::kFloat8_e4m3fn, "gemv_allreduce: int8 or float8_e4m3fn weights");
    TORCH_CHECK(scale->scalar_type() == at::kFloat && scale->is_contiguous() && (scale->numel() == 1 || scale->numel() == w.size(0)));
  } else {
    check_gemv_inputs(x, w);
  }
  const int world = recv_ptrs.size();
  TORCH_CHECK(world >= 2 && world <= SYMM_MAX_RANKS);
  TORCH_CHECK(step.is_cuda() && step.scalar_type() == at::kInt && step.numel() >= 1, "gemv_allreduce: step counter must be a CUDA int32 tensor");
  TORCH_CHECK(gemv2_supported(x.size(0), x.size(1), wt), "gemv_allreduce: activations too wide for shared memory");
  const int N = w.size(0);
  TORCH_CHECK(N <= n_max && (N + 15) / 16 <= SYMM_MAX_TILES, "gemv_allreduce: output too wide for the workspace");
  c10::cuda::CUDAGuard guard(x.device());
  bool x_in_smem;
  at::Tensor xn = prenorm_if_needed(x, c10::nullopt, 0, 0, x_in_smem);
  auto y = at::empty({x.size(0), N}, x.options());
  if (residual.has_value())
    TORCH_CHECK(residual->is_contiguous() && residual->size(0) == x.size(0) && residual->size(1) == N && is_bf16(*residual));
  GemvParams p{};
  fill_params(p, xn, w, bias, c10::nullopt, 0, 0, 0, residual, scale, y, x_in_smem);
  for (int i = 0; i < world; ++i) p.symm.recv[i] = reinterpret_cast<float*>(recv_ptrs[i]);
  p.symm.step = reinterpret_cast<const uint32_t*>(step.data_ptr<int>());
  p.symm.rank = rank;
  p.symm.world = world;
  p.symm.parity = parity;
  p.symm.call = call;
  p.symm.n_max = n_max;
  run_gemv(p, 1, x.device());
  return y;
}

// Routed experts of a decode step: x [T,H], w_gate_up [E,2I,H], w_down [E,H,I], topk_w/topk_i [T,k] -> [T,H] (local experts only;
// the caller all-reduces across EP/TP ranks).
at::Tensor moe_decode(const at::Tensor& x, const at::Tensor& w_gate_up, const at::Tensor& w_down, const at::Tensor& topk_w,
                      const at::Tensor& topk_i, int64_t expert_offset) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous() && is_bf16(x) && is_bf16(w_gate_up) && is_bf16(w_down));
  TORCH_CHECK(w_gate_up.dim() == 3 && w_down.dim() == 3 && w_gate_up.is_contiguous() && w_down.is_contiguous());
  const int T = x.size(0), H = x.size(1), E = w_gate_up.size(0), I = w_gate_up.size(1) / 2, k = topk_i.size(1);
  TORCH_CHECK(w_gate_up.size(2) == H && w_down.size(0) == E && w_down.size(1) == H && w_down.size(2) == I);
  TORCH_CHECK(topk_w.scalar_type() == at::kFloat && topk_i.scalar_type() == at::kInt && topk_w.is_contiguous() &&
              topk_i.is_contiguous() && topk_w.numel() == (int64_t)T * k && topk_i.size(0) == T);
  c10::cuda::CUDAGuard guard(x.device());
  auto u = at::empty({(int64_t)T * k, I}, x.options());
  auto y = at::zeros({T, H}, x.options().dtype(at::kFloat));
  static int n_sms = 0;
  if (n_sms == 0) n_sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  moe_decode_launch(x.data_ptr(), w_gate_up.data_ptr(), w_down.data_ptr(), topk_w.data_ptr<float>(), topk_i.data_ptr<int>(),
                    u.data_ptr(), y.data_ptr<float>(), T, k, H, I, E, (int)expert_offset, n_sms, cur_stream());
  return y.to(x.scalar_type());
}

// int8 / fp8 weight [N, K] x scale -> bf16 into `out` (a caller-owned scratch: one buffer serves every layer of a model)
at::Tensor dequant_bf16(const at::Tensor& w, const at::Tensor& scale, at::Tensor out) {
  TORCH_CHECK(w.is_cuda() && w.dim() == 2 && w.is_contiguous() && scale.scalar_type() == at::kFloat && scale.is_contiguous());
  const bool i8 = w.scalar_type() == at::kChar;
  TORCH_CHECK(i8 || w.scalar_type() == at::kFloat8_e4m3fn, "dequant_bf16: int8 or float8_e4m3fn weights");
  TORCH_CHECK(scale.numel() == 1 || scale.numel() == w.size(0));
  TORCH_CHECK(out.is_cuda() && is_bf16(out) && out.is_contiguous() && out.numel() >= w.numel());
  c10::cuda::CUDAGuard guard(w.device());
  dequant_bf16_launch(w.data_ptr(), i8 ? 1 : 2, scale.data_ptr<float>(), (int)scale.numel(), out.data_ptr(), (int)w.size(0), (int)w.size(1),
                      cur_stream());
  return out.flatten().narrow(0, 0, w.numel()).view({w.size(0), w.size(1)});
}

// Routed experts of a prefill-sized batch (any N): permutation on the device + two grouped tcgen05 GEMMs + weighted combine.
// act: 1 silu*up, 2 gelu_tanh*up, 3 gelu*up.  gate_up_bias [E, 2I] / down_bias [E, H] optional (bf16).
at::Tensor moe_grouped(const at::Tensor& x, const at::Tensor& w_gate_up, const at::Tensor& w_down, const at::Tensor& topk_w,
                       const at::Tensor& topk_i, int64_t expert_offset, int64_t act, bool scale_input,
                       const c10::optional<at::Tensor>& gate_up_bias, const c10::optional<at::Tensor>& down_bias,
                       const c10::optional<at::Tensor>& gate_up_scale, const c10::optional<at::Tensor>& down_scale) {
  // W8A8 experts: fp8-e4m3 weights with per-(expert, channel) scales; the permuted activations and the intermediate are quantised
  // per row (quant.cu) and both grouped GEMMs run on the kind::f8f6f4 tensor-core path
  const bool fp8 = gate_up_scale.has_value();
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous() && is_bf16(x));
  if (fp8) {
    TORCH_CHECK(w_gate_up.scalar_type() == at::kFloat8_e4m3fn && w_down.scalar_type() == at::kFloat8_e4m3fn && down_scale.has_value());
    TORCH_CHECK(gate_up_scale->scalar_type() == at::kFloat && gate_up_scale->is_contiguous() && down_scale->scalar_type() == at::kFloat &&
                down_scale->is_contiguous());
  } else {
    TORCH_CHECK(is_bf16(w_gate_up) && is_bf16(w_down));
  }
  TORCH_CHECK(w_gate_up.dim() == 3 && w_down.dim() == 3 && w_gate_up.is_contiguous() && w_down.is_contiguous());
  const int N = x.size(0), H = x.size(1), E = w_gate_up.size(0), I = w_gate_up.size(1) / 2, k = topk_i.size(1);
  TORCH_CHECK(w_gate_up.size(2) == H && w_down.size(0) == E && w_down.size(1) == H && w_down.size(2) == I);
  TORCH_CHECK(H % 64 == 0 && I % 64 == 0, "moe_grouped: hidden and intermediate sizes must be multiples of 64");
  if (fp8) {
    TORCH_CHECK(H % 128 == 0 && I % 128 == 0 && H <= 16384 && I <= 16384, "moe_grouped (fp8): sizes must be multiples of 128, <= 16384");
    TORCH_CHECK(gate_up_scale->numel() == (int64_t)E * 2 * I && down_scale->numel() == (int64_t)E * H);
  }
  TORCH_CHECK(topk_w.scalar_type() == at::kFloat && topk_i.scalar_type() == at::kInt && topk_w.is_contiguous() &&
              topk_i.is_contiguous() && topk_w.numel() == (int64_t)N * k && topk_i.size(0) == N);
  if (gate_up_bias) TORCH_CHECK(is_bf16(*gate_up_bias) && gate_up_bias->is_contiguous() && gate_up_bias->numel() == (int64_t)E * 2 * I);
  if (down_bias) TORCH_CHECK(is_bf16(*down_bias) && down_bias->is_contiguous() && down_bias->numel() == (int64_t)E * H);
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t R = (((int64_t)N * k + (int64_t)E * 127) + 127) / 128 * 128;   // static bound: graph-stable shapes
  auto iopt = x.options().dtype(at::kInt);
  auto pos = at::empty({(int64_t)N * k}, iopt);
  auto row_entry = at::empty({R}, iopt);
  auto tile_expert = at::empty({R / 128}, iopt);
  auto xp = at::empty({R, H}, x.options());
  auto h = at::empty({R, I}, x.options());
  auto y = at::empty({R, H}, x.options());
  auto out = at::empty({N, H}, x.options());
  auto st = cur_stream();
  moe_plan_launch(topk_i.data_ptr<int>(), N * k, (int)expert_offset, E, (int)R, pos.data_ptr<int>(), row_entry.data_ptr<int>(),
                  tile_expert.data_ptr<int>(), st);
  moe_gather_launch(x.data_ptr(), topk_w.data_ptr<float>(), row_entry.data_ptr<int>(), tile_expert.data_ptr<int>(), xp.data_ptr(), (int)R,
                    H, k, scale_input ? 1 : 0, st);
  const float inf = std::numeric_limits<float>::infinity();
  if (fp8) {
    auto bopt = x.options().dtype(at::kByte);
    auto xq = at::empty({R, H}, bopt), hq = at::empty({R, I}, bopt);
    auto xs = at::empty({R}, x.options().dtype(at::kFloat)), hs = at::empty({R}, x.options().dtype(at::kFloat));
    rmsnorm_quant_launch(xp.data_ptr(), nullptr, xq.data_ptr(), xs.data_ptr<float>(), (int)R, H, 0.f, 0.f, inf, st);
    gemm_grouped_launch(xq.data_ptr(), w_gate_up.data_ptr(), gate_up_bias ? gate_up_bias->data_ptr() : nullptr, h.data_ptr(), (int)R, 2 * I,
                        H, E, (int)act, tile_expert.data_ptr<int>(), xs.data_ptr<float>(), gate_up_scale->data_ptr<float>(), st);
    rmsnorm_quant_launch(h.data_ptr(), nullptr, hq.data_ptr(), hs.data_ptr<float>(), (int)R, I, 0.f, 0.f, inf, st);
    gemm_grouped_launch(hq.data_ptr(), w_down.data_ptr(), down_bias ? down_bias->data_ptr() : nullptr, y.data_ptr(), (int)R, H, I, E, 0,
                        tile_expert.data_ptr<int>(), hs.data_ptr<float>(), down_scale->data_ptr<float>(), st);
  } else {
    gemm_grouped_launch(xp.data_ptr(), w_gate_up.data_ptr(), gate_up_bias ? gate_up_bias->data_ptr() : nullptr, h.data_ptr(), (int)R,
                        2 * I, H, E, (int)act, tile_expert.data_ptr<int>(), nullptr, nullptr, st);
    gemm_grouped_launch(h.data_ptr(), w_down.data_ptr(), down_bias ? down_bias->data_ptr() : nullptr, y.data_ptr(), (int)R, H, I, E, 0,
                        tile_expert.data_ptr<int>(), nullptr, nullptr, st);
  }
  moe_combine_launch(y.data_ptr(), topk_w.data_ptr<float>(), pos.data_ptr<int>(), out.data_ptr(), N, H, k, scale_input ? 1 : 0, st);
  return out;
}

// ---- symmetric (peer-mapped) workspace ---------------------------------------------------------------------
std::tuple<int64_t, pybind11::bytes> symm_alloc(int64_t nbytes) {
  void* p = nullptr;
  NXDI_CUDA_OK(cudaMalloc(&p, nbytes));
  NXDI_CUDA_OK(cudaMemset(p, 0, nbytes));
  cudaIpcMemHandle_t h;
  NXDI_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  NXDI_CUDA_OK(cudaDeviceSynchronize());
  return {reinterpret_cast<int64_t>(p), pybind11::bytes(reinterpret_cast<const char*>(&h), sizeof(h))};
}
int64_t symm_open(const std::string& handle) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  NXDI_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return reinterpret_cast<int64_t>(p);
}
void symm_close(int64_t ptr) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)); }
void symm_free(int64_t ptr) { cudaFree(reinterpret_cast<void*>(ptr)); }
at::Tensor symm_as_tensor(int64_t ptr, int64_t nbytes, int64_t device) {
  auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device);
  return at::from_blob(reinterpret_cast<void*>(ptr), {nbytes}, opts);
}

// ---- rope / cache appends -------------------------------------------------------------------------------------
at::Tensor rope_kv_append(const at::Tensor& qkv, const at::Tensor& cos, const at::Tensor& sin, at::Tensor& k_cache,
                          at::Tensor& v_cache, const at::Tensor& lines, const at::Tensor& positions, int64_t nq, int64_t nkv,
                          int64_t D, const c10::optional<at::Tensor>& q_norm, const c10::optional<at::Tensor>& k_norm,
                          double eps) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 3 && qkv.is_contiguous() && is_bf16(qkv) && is_bf16(k_cache));
  const int B = qkv.size(0), T = qkv.size(1);
  TORCH_CHECK(qkv.size(2) == (nq + 2 * nkv) * D);
  TORCH_CHECK(cos.scalar_type() == at::kFloat && cos.is_contiguous() && sin.is_contiguous() &&
              cos.numel() == (int64_t)B * T * D / 2 && sin.numel() == cos.numel());
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.size(1) == nkv && k_cache.size(3) == D && k_cache.is_contiguous() &&
              v_cache.is_contiguous());
  TORCH_CHECK(lines.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && positions.is_contiguous() &&
              lines.numel() == B && positions.numel() == B * T);
  c10::cuda::CUDAGuard guard(qkv.device());
  auto q = at::empty({B, T, nq, D}, qkv.options());
  rope_kv_append_launch(qkv.data_ptr(), cos.data_ptr<float>(), sin.data_ptr<float>(), q.data_ptr(), k_cache.data_ptr(),
                        v_cache.data_ptr(), lines.data_ptr<int>(), positions.data_ptr<int>(), optr(q_norm), optr(k_norm),
                        (float)eps, B, T, (int)nq, (int)nkv, (int)D, (int)k_cache.size(0), (int)k_cache.size(2), cur_stream());
  return q;
}

// prefill flavour: the same single pass also returns the rotated k and the v of the new tokens as contiguous [B, T, nkv, D]
std::tuple<at::Tensor, at::Tensor, at::Tensor> rope_kv_split_append(const at::Tensor& qkv, const at::Tensor& cos, const at::Tensor& sin,
                                                                   at::Tensor& k_cache, at::Tensor& v_cache, const at::Tensor& lines,
                                                                   const at::Tensor& positions, int64_t nq, int64_t nkv, int64_t D,
                                                                   const c10::optional<at::Tensor>& q_norm,
                                                                   const c10::optional<at::Tensor>& k_norm, double eps) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 3 && qkv.is_contiguous() && is_bf16(qkv) && is_bf16(k_cache));
  const int B = qkv.size(0), T = qkv.size(1);
  TORCH_CHECK(qkv.size(2) == (nq + 2 * nkv) * D);
  TORCH_CHECK(cos.scalar_type() == at::kFloat && cos.is_contiguous() && sin.is_contiguous() &&
              cos.numel() == (int64_t)B * T * D / 2 && sin.numel() == cos.numel());
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.size(1) == nkv && k_cache.size(3) == D && k_cache.is_contiguous() &&
              v_cache.is_contiguous());
  TORCH_CHECK(lines.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && positions.is_contiguous() &&
              lines.numel() == B && positions.numel() == B * T);
  c10::cuda::CUDAGuard guard(qkv.device());
  auto q = at::empty({B, T, nq, D}, qkv.options());
  auto k = at::empty({B, T, nkv, D}, qkv.options());
  auto v = at::empty({B, T, nkv, D}, qkv.options());
  rope_kv_append_launch(qkv.data_ptr(), cos.data_ptr<float>(), sin.data_ptr<float>(), q.data_ptr(), k_cache.data_ptr(),
                        v_cache.data_ptr(), lines.data_ptr<int>(), positions.data_ptr<int>(), optr(q_norm), optr(k_norm),
                        (float)eps, B, T, (int)nq, (int)nkv, (int)D, (int)k_cache.size(0), (int)k_cache.size(2), cur_stream(),
                        k.data_ptr(), v.data_ptr());
  return {q, k, v};
}

void kv_append(at::Tensor& k_cache, at::Tensor& v_cache, const at::Tensor& k_new, const at::Tensor& v_new,
               const at::Tensor& lines, const at::Tensor& positions) {
  TORCH_CHECK(k_new.is_cuda() && k_new.dim() == 4 && k_new.is_contiguous() && v_new.is_contiguous());
  TORCH_CHECK(k_cache.is_contiguous() && v_cache.is_contiguous() && k_cache.scalar_type() == k_new.scalar_type());
  const int B = k_new.size(0), T = k_new.size(1), H = k_new.size(2), D = k_new.size(3);
  TORCH_CHECK((D * k_new.element_size()) % 16 == 0 && v_new.size(3) == D && k_cache.size(1) == H && k_cache.size(3) == D);
  TORCH_CHECK(lines.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && positions.is_contiguous());
  c10::cuda::CUDAGuard guard(k_new.device());
  kv_append_launch(k_new.data_ptr(), v_new.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), lines.data_ptr<int>(),
                   positions.data_ptr<int>(), B, T, H, D * (int)k_new.element_size(), (int)k_cache.size(0),
                   (int)k_cache.size(2), cur_stream());
}

void paged_kv_append(at::Tensor& k_cache, at::Tensor& v_cache, const at::Tensor& k_new, const at::Tensor& v_new,
                     const at::Tensor& slots) {
  TORCH_CHECK(k_new.is_cuda() && k_new.is_contiguous() && v_new.is_contiguous() && k_cache.is_contiguous() &&
              v_cache.is_contiguous() && k_cache.dim() == 4);
  const int H = k_cache.size(2), D = k_cache.size(3);
  const int ntok = k_new.numel() / (H * D);
  TORCH_CHECK(slots.numel() == ntok && slots.scalar_type() == at::kInt && slots.is_contiguous());
  c10::cuda::CUDAGuard guard(k_new.device());
  paged_kv_append_launch(k_new.data_ptr(), v_new.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), slots.data_ptr<int>(),
                         ntok, H * D * (int)k_new.element_size(), (int)(k_cache.size(0) * k_cache.size(1)), cur_stream());
}

// ---- sampling ---------------------------------------------------------------------------------------------------
static Scratch& scratch(const at::Device& dev, int64_t nf, int64_t ni, int64_t nt) {
  static std::map<int, Scratch> all;
  auto& s = all[dev.index()];
  auto o = at::TensorOptions().device(dev);
  // generous first sizes (16 MB of partials covers B = 8, 8 kv heads, 32 splits at head_dim 128) so that growth is rare
  s.grow(s.f, std::max<int64_t>(nf, 4 << 20), o.dtype(at::kFloat), false);
  s.grow(s.i, std::max<int64_t>(ni, 1 << 16), o.dtype(at::kInt), false);
  s.grow(s.tickets, std::max<int64_t>(nt, 1 << 14), o.dtype(at::kInt), true);
  return s;
}

at::Tensor argmax(const at::Tensor& logits, const std::vector<int64_t>& slot_ptrs, const c10::optional<at::Tensor>& step,
                  int64_t rank, int64_t parity, int64_t call, int64_t rows_max) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  TORCH_CHECK(logits.scalar_type() == at::kFloat || is_bf16(logits), "argmax: fp32 or bf16 logits");
  const int B = logits.size(0), V = logits.size(1);
  c10::cuda::CUDAGuard guard(logits.device());
  auto out = at::empty({B}, logits.options().dtype(at::kLong));
  const int nsplit = std::max(1, std::min(64, V / 2048));
  auto& ws = scratch(logits.device(), (int64_t)B * nsplit, (int64_t)B * nsplit, B);
  ArgmaxSymm sy{};
  const int world = slot_ptrs.size();
  if (world > 1) {
    TORCH_CHECK(world <= SYMM_MAX_RANKS && step.has_value() && step->is_cuda() && step->scalar_type() == at::kInt && B <= rows_max,
                "argmax exchange: bad symmetric arguments");
    for (int i = 0; i < world; ++i) sy.slots[i] = reinterpret_cast<float*>(slot_ptrs[i]);
    sy.step = reinterpret_cast<const uint32_t*>(step->data_ptr<int>());
    sy.rank = rank; sy.world = world; sy.parity = parity; sy.call = call; sy.rows_max = rows_max;
  }
  argmax_launch(logits.data_ptr(), is_bf16(logits) ? 1 : 0, out.data_ptr<int64_t>(), ws.f.data_ptr<float>(), ws.i.data_ptr<int>(),
                reinterpret_cast<unsigned*>(ws.tickets.data_ptr<int>()), B, V, (int)logits.stride(0), nsplit,
                world > 1 ? &sy : nullptr, cur_stream());
  return out;
}

at::Tensor topk_sample(const at::Tensor& logits, const at::Tensor& top_k, const at::Tensor& top_p,
                       const at::Tensor& temperature, const at::Tensor& rand, int64_t global_topk) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  TORCH_CHECK(logits.scalar_type() == at::kFloat || is_bf16(logits), "topk_sample: fp32 or bf16 logits");
  TORCH_CHECK(global_topk >= 1 && global_topk <= 256);
  const int B = logits.size(0), V = logits.size(1);
  TORCH_CHECK(top_k.scalar_type() == at::kInt && top_k.numel() == B && top_p.scalar_type() == at::kFloat &&
              temperature.scalar_type() == at::kFloat && rand.scalar_type() == at::kFloat && rand.numel() == B);
  c10::cuda::CUDAGuard guard(logits.device());
  auto out = at::empty({B}, logits.options().dtype(at::kLong));
  topk_sample_launch(logits.data_ptr(), is_bf16(logits) ? 1 : 0, top_k.data_ptr<int>(), top_p.data_ptr<float>(),
                     temperature.data_ptr<float>(), rand.data_ptr<float>(), out.data_ptr<int64_t>(), B, V, (int)logits.stride(0),
                     (int)global_topk, cur_stream());
  return out;
}

// ---- attention -----------------------------------------------------------------------------------------------------
int pick_nsplit(int B, int Hkv, int S_hint) {
  const int ctas = B * Hkv;
  // short contexts: one 64-key tile per CTA (latency bound: tiles in parallel, not in sequence);
  // long contexts: ~4 tiles per CTA once the chip is full
  // (measured: splitting a <=512-key context costs more in the combine than the 4-deep tile prefetch hides)
  int want = S_hint <= 512 ? 1 : (S_hint + 255) / 256;
  int cap = std::max(1, (2 * 148) / std::max(ctas, 1));  // fill the chip, not more
  return std::max(1, std::min({want, cap, 64}));
}

at::Tensor attention_decode(const at::Tensor& q, const at::Tensor& k_cache, const at::Tensor& v_cache, const at::Tensor& lines,
                            const at::Tensor& positions, double scale, int64_t window, const c10::optional<at::Tensor>& sinks,
                            int64_t s_hint) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 4 && q.is_contiguous() && is_bf16(q) && is_bf16(k_cache));
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.is_contiguous() && v_cache.is_contiguous());
  const int B = q.size(0), T = q.size(1), Hq = q.size(2), D = q.size(3);
  const int L = k_cache.size(0), Hkv = k_cache.size(1), S = k_cache.size(2);
  TORCH_CHECK(k_cache.size(3) == D && Hq % Hkv == 0);
  TORCH_CHECK(lines.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && positions.is_contiguous() &&
              lines.numel() == B && positions.numel() == B * T);
  c10::cuda::CUDAGuard guard(q.device());
  auto out = at::empty_like(q);
  AttnDecodeParams p{};
  p.q = q.data_ptr(); p.k_cache = k_cache.data_ptr(); p.v_cache = v_cache.data_ptr(); p.out = out.data_ptr();
  p.lines = lines.data_ptr<int>(); p.positions = positions.data_ptr<int>(); p.block_table = nullptr;
  p.sinks = sinks.has_value() ? sinks->data_ptr<float>() : nullptr;
  p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.D = D; p.S = S; p.L = L; p.window = window; p.scale = (float)scale;
  p.nsplit = pick_nsplit(B, Hkv, s_hint > 0 ? (int)std::min<int64_t>(s_hint, S) : S);
  auto& ws = scratch(q.device(), (int64_t)B * Hkv * p.nsplit * 64 * (D + 2), 0, (int64_t)B * Hkv);
  p.ws_o = ws.f.data_ptr<float>();
  p.ws_ml = p.ws_o + (int64_t)B * Hkv * p.nsplit * 64 * D;
  p.tickets = reinterpret_cast<unsigned*>(ws.tickets.data_ptr<int>());
  attention_decode_launch(p, cur_stream());
  return out;
}

// q/k RMSNorm + RoPE + KV-cache append + split-KV flash decode in ONE kernel (the prologue of the attention kernel).
at::Tensor rope_attention_decode(const at::Tensor& qkv, const at::Tensor& cos, const at::Tensor& sin, at::Tensor& k_cache,
                                 at::Tensor& v_cache, const at::Tensor& lines, const at::Tensor& write_pos,
                                 const at::Tensor& positions, int64_t nq, int64_t nkv, int64_t D, double scale, int64_t window,
                                 const c10::optional<at::Tensor>& sinks, const c10::optional<at::Tensor>& q_norm,
                                 const c10::optional<at::Tensor>& k_norm, double eps, int64_t s_hint) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 3 && qkv.is_contiguous() && is_bf16(qkv) && is_bf16(k_cache));
  const int B = qkv.size(0), T = qkv.size(1);
  TORCH_CHECK(qkv.size(2) == (nq + 2 * nkv) * D && (D == 64 || D == 128) && nq % nkv == 0);
  TORCH_CHECK(cos.scalar_type() == at::kFloat && cos.is_contiguous() && sin.is_contiguous() &&
              cos.numel() == (int64_t)B * T * D / 2 && sin.numel() == cos.numel());
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.size(1) == nkv && k_cache.size(3) == D && k_cache.is_contiguous() &&
              v_cache.is_contiguous());
  TORCH_CHECK(lines.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && write_pos.scalar_type() == at::kInt &&
              positions.is_contiguous() && write_pos.is_contiguous() && lines.numel() == B && positions.numel() == B * T &&
              write_pos.numel() == B * T);
  c10::cuda::CUDAGuard guard(qkv.device());
  auto out = at::empty({B, T, nq, D}, qkv.options());
  const int L = k_cache.size(0), S = k_cache.size(2);
  AttnDecodeParams p{};
  p.q = nullptr; p.k_cache = k_cache.data_ptr(); p.v_cache = v_cache.data_ptr(); p.out = out.data_ptr();
  p.lines = lines.data_ptr<int>(); p.positions = positions.data_ptr<int>(); p.block_table = nullptr;
  p.sinks = sinks.has_value() ? sinks->data_ptr<float>() : nullptr;
  p.B = B; p.T = T; p.Hq = nq; p.Hkv = nkv; p.D = D; p.S = S; p.L = L; p.window = window; p.scale = (float)scale;
  p.qkv = qkv.data_ptr(); p.cos = cos.data_ptr<float>(); p.sin = sin.data_ptr<float>();
  p.q_norm = optr(q_norm); p.k_norm = optr(k_norm); p.write_pos = write_pos.data_ptr<int>(); p.norm_eps = (float)eps;
  p.nsplit = pick_nsplit(B, nkv, s_hint > 0 ? (int)std::min<int64_t>(s_hint, S) : S);
  auto& ws = scratch(qkv.device(), (int64_t)B * nkv * p.nsplit * 64 * (D + 2), 0, (int64_t)B * nkv);
  p.ws_o = ws.f.data_ptr<float>();
  p.ws_ml = p.ws_o + (int64_t)B * nkv * p.nsplit * 64 * D;
  p.tickets = reinterpret_cast<unsigned*>(ws.tickets.data_ptr<int>());
  attention_decode_launch(p, cur_stream());
  return out;
}

at::Tensor paged_attention_decode(const at::Tensor& q, const at::Tensor& k_cache, const at::Tensor& v_cache,
                                  const at::Tensor& block_table, const at::Tensor& positions, double scale, int64_t window,
                                  const c10::optional<at::Tensor>& sinks) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 4 && q.is_contiguous() && is_bf16(q) && is_bf16(k_cache));
  TORCH_CHECK(k_cache.dim() == 4 && k_cache.is_contiguous() && v_cache.is_contiguous());
  const int B = q.size(0), T = q.size(1), Hq = q.size(2), D = q.size(3);
  const int bs = k_cache.size(1), Hkv = k_cache.size(2);
  TORCH_CHECK(k_cache.size(3) == D && Hq % Hkv == 0 && block_table.dim() == 2 && block_table.size(0) == B &&
              block_table.scalar_type() == at::kInt && block_table.is_contiguous());
  TORCH_CHECK(positions.scalar_type() == at::kInt && positions.is_contiguous() && positions.numel() == B * T);
  c10::cuda::CUDAGuard guard(q.device());
  auto out = at::empty_like(q);
  AttnDecodeParams p{};
  p.q = q.data_ptr(); p.k_cache = k_cache.data_ptr(); p.v_cache = v_cache.data_ptr(); p.out = out.data_ptr();
  p.lines = nullptr; p.positions = positions.data_ptr<int>(); p.block_table = block_table.data_ptr<int>();
  p.sinks = sinks.has_value() ? sinks->data_ptr<float>() : nullptr;
  p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.D = D; p.S = 0; p.L = 0; p.window = window; p.scale = (float)scale;
  p.block_size = bs; p.max_blocks = block_table.size(1);
  p.nsplit = pick_nsplit(B, Hkv, p.max_blocks * bs);
  auto& ws = scratch(q.device(), (int64_t)B * Hkv * p.nsplit * 64 * (D + 2), 0, (int64_t)B * Hkv);
  p.ws_o = ws.f.data_ptr<float>();
  p.ws_ml = p.ws_o + (int64_t)B * Hkv * p.nsplit * 64 * D;
  p.tickets = reinterpret_cast<unsigned*>(ws.tickets.data_ptr<int>());
  attention_decode_launch(p, cur_stream());
  return out;
}

at::Tensor attention_prefill(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, double scale, int64_t window,
                             const c10::optional<at::Tensor>& sinks, bool causal) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 4 && q.is_contiguous() && k.is_contiguous() && v.is_contiguous() && is_bf16(q) &&
              is_bf16(k) && is_bf16(v));
  const int B = q.size(0), T = q.size(1), Hq = q.size(2), D = q.size(3), Hkv = k.size(2);
  TORCH_CHECK(k.size(0) == B && k.size(1) == T && k.size(3) == D && v.sizes() == k.sizes() && Hq % Hkv == 0);
  c10::cuda::CUDAGuard guard(q.device());
  auto out = at::empty_like(q);
  AttnPrefillParams p{};
  p.q = q.data_ptr(); p.k = k.data_ptr(); p.v = v.data_ptr(); p.out = out.data_ptr();
  p.sinks = sinks.has_value() ? sinks->data_ptr<float>() : nullptr;
  p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.D = D; p.window = window; p.scale = (float)scale;
  p.causal = causal ? 1 : 0;
  attention_prefill_launch(p, cur_stream());
  return out;
}

}  // namespace nxdi

namespace nxdi {
// ---- W8A8 fp8 ------------------------------------------------------------------------------------------------------------------
std::tuple<at::Tensor, at::Tensor> rmsnorm_quant(const at::Tensor& x, const c10::optional<at::Tensor>& gamma, double eps, double offset,
                                                 double clamp) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous() && is_bf16(x));
  if (gamma.has_value()) TORCH_CHECK(is_bf16(*gamma) && gamma->numel() == x.size(1));
  c10::cuda::CUDAGuard guard(x.device());
  auto q = at::empty(x.sizes(), x.options().dtype(at::kFloat8_e4m3fn));
  auto scale = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  rmsnorm_quant_launch(x.data_ptr(), optr(gamma), q.data_ptr(), scale.data_ptr<float>(), (int)x.size(0), (int)x.size(1), (float)eps,
                       (float)offset, (float)clamp, cur_stream());
  return {q, scale};
}
at::Tensor gemm_fp8(const at::Tensor& xq, const at::Tensor& a_scale, const at::Tensor& w, const at::Tensor& w_scale,
                    const c10::optional<at::Tensor>& bias, int64_t act, const c10::optional<at::Tensor>& residual) {
  TORCH_CHECK(xq.is_cuda() && xq.dim() == 2 && xq.is_contiguous() && xq.scalar_type() == at::kFloat8_e4m3fn);
  TORCH_CHECK(w.dim() == 2 && w.is_contiguous() && w.scalar_type() == at::kFloat8_e4m3fn && w.size(1) == xq.size(1));
  const int M = xq.size(0), K = xq.size(1), N = w.size(0);
  TORCH_CHECK(K % 128 == 0 && a_scale.scalar_type() == at::kFloat && a_scale.numel() == M && w_scale.scalar_type() == at::kFloat &&
              (w_scale.numel() == 1 || w_scale.numel() == N) && a_scale.is_contiguous() && w_scale.is_contiguous());
  const bool glu = act != 0;
  const int n_out = glu ? N / 2 : N;
  TORCH_CHECK(n_out % 8 == 0);
  c10::cuda::CUDAGuard guard(xq.device());
  auto y = at::empty({M, n_out}, xq.options().dtype(at::kBFloat16));
  if (residual.has_value()) TORCH_CHECK(!glu && is_bf16(*residual) && residual->is_contiguous() && residual->numel() == y.numel());
  gemm_fp8_launch(xq.data_ptr(), K, w.data_ptr(), a_scale.data_ptr<float>(), w_scale.data_ptr<float>(), (int)w_scale.numel(), optr(bias),
                  optr(residual), y.data_ptr(), n_out, M, N, K, (int)act, cur_stream());
  return y;
}

// Row-parallel GEMM with the reduce-scatter fused into the kernel: x [M, K] (M = segs * rows_per_seg, rows of a segment split
// evenly over the ranks) -> private output [M / world, N] = sum over ranks of x_r w_r^T (+ bias on rank 0) (+ residual).
// `staging` is this rank's symmetric buffer for the partial sums, `mc_ptr` its multicast address.
at::Tensor gemm_reduce_scatter(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& bias, at::Tensor& staging,
                               int64_t mc_ptr, const std::vector<int64_t>& flag_ptrs, int64_t max_tiles, const at::Tensor& step,
                               int64_t call, int64_t rank, int64_t rows_per_seg, const c10::optional<at::Tensor>& residual,
                               int64_t bcast_mc_ptr, const std::vector<int64_t>& done_ptrs, const c10::optional<at::Tensor>& cta_counter) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && is_bf16(x) && is_bf16(w) && w.is_contiguous());
  const int M = x.size(0), K = x.size(1), N = w.size(0), world = flag_ptrs.size();
  TORCH_CHECK(K % 64 == 0 && x.stride(1) == 1 && x.stride(0) % 8 == 0 && world >= 2 && world <= SYMM_MAX_RANKS);
  TORCH_CHECK(is_bf16(staging) && staging.is_contiguous() && staging.numel() == (int64_t)M * N);
  c10::cuda::CUDAGuard guard(x.device());
  const bool bcast = bcast_mc_ptr != 0;   // all-reduce flavour: result lands in the symmetric output buffer of every rank
  auto out = bcast ? at::empty({0}, x.options()) : at::empty({M / world, N}, x.options());
  if (residual.has_value())
    TORCH_CHECK(is_bf16(*residual) && residual->is_contiguous() && residual->numel() == (bcast ? (int64_t)M * N : out.numel()));
  if (bcast) TORCH_CHECK((int)done_ptrs.size() == world && cta_counter.has_value() && cta_counter->is_cuda() && cta_counter->scalar_type() == at::kInt);
  GemmRsArgs rs{};
  for (int i = 0; i < world; ++i) rs.flag_ptrs[i] = flag_ptrs[i];
  rs.mc = reinterpret_cast<const void*>(mc_ptr);
  rs.out = bcast ? nullptr : out.data_ptr();
  rs.bcast_mc = reinterpret_cast<const void*>(bcast_mc_ptr);
  for (int i = 0; i < world && bcast; ++i) rs.done_ptrs[i] = done_ptrs[i];
  rs.cta_counter = bcast ? cta_counter->data_ptr() : nullptr;
  rs.residual = residual.has_value() ? residual->data_ptr() : nullptr;
  rs.step = step.data_ptr();
  rs.call = (int)call; rs.rank = (int)rank; rs.world = world; rs.rows_per_seg = (int)rows_per_seg; rs.max_tiles = (int)max_tiles;
  gemm_tcgen05_launch(x.data_ptr(), (int)x.stride(0), w.data_ptr(), optr(bias), nullptr, staging.data_ptr(), N, M, N, K, 0, cur_stream(), &rs);
  return out;
}

at::Tensor attention_prefill_tc(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, double scale, int64_t window,
                                const c10::optional<at::Tensor>& sinks, bool causal, double softcap) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 4 && q.is_contiguous() && k.is_contiguous() && v.is_contiguous() && is_bf16(q) && is_bf16(k) && is_bf16(v));
  const int B = q.size(0), T = q.size(1), Hq = q.size(2), D = q.size(3), Hkv = k.size(2);
  TORCH_CHECK(D == 128 && k.size(1) == T && k.size(3) == D && v.sizes() == k.sizes() && Hq % Hkv == 0, "attention_prefill_tc: head_dim 128, equal q/k lengths");
  c10::cuda::CUDAGuard guard(q.device());
  auto out = at::empty_like(q);
  attention_prefill_tc_launch(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), sinks.has_value() ? sinks->data_ptr<float>() : nullptr,
                              B, T, Hq, Hkv, (float)scale, causal ? 1 : 0, (int)window, (float)softcap, cur_stream());
  return out;
}

// ---- symmetric heap + NVLS collectives -------------------------------------------------------------------------------------
// mode 0 all-reduce (in place in the symmetric buffer; `out` = private copy (+ residual) when given), 1 reduce-scatter, 2 all-gather
void nvls_collective(int64_t mode, const std::vector<int64_t>& sig_ptrs, const at::Tensor& step, int64_t rank, int64_t call,
                     int64_t mc_ptr, int64_t local_ptr, const c10::optional<at::Tensor>& residual, const c10::optional<at::Tensor>& out,
                     int64_t segs, int64_t rows_per_seg, int64_t row_elems) {
  TORCH_CHECK(step.is_cuda() && step.scalar_type() == at::kInt);
  if (residual.has_value()) TORCH_CHECK(is_bf16(*residual) && residual->is_contiguous());
  if (out.has_value()) TORCH_CHECK(is_bf16(*out) && out->is_contiguous());
  c10::cuda::CUDAGuard guard(step.device());
  std::vector<long long> sp(sig_ptrs.begin(), sig_ptrs.end());
  nvls_collective_launch((int)mode, sp.data(), step.data_ptr(), (int)rank, (int)sp.size(), (int)call, reinterpret_cast<void*>(mc_ptr),
                         reinterpret_cast<void*>(local_ptr), residual.has_value() ? residual->data_ptr() : nullptr,
                         out.has_value() ? out->data_ptr() : nullptr, (int)segs, (int)rows_per_seg, (int)row_elems, cur_stream());
}

// ---- persistent decode-step kernel -------------------------------------------------------------------------------------
int64_t dstep_new_b(int64_t T) { return dstep_new((int)T); }
void dstep_set_symm_b(int64_t h, const std::vector<int64_t>& recv_ptrs, const at::Tensor& step, int64_t rank, int64_t n_max) {
  std::vector<long long> v(recv_ptrs.begin(), recv_ptrs.end());
  dstep_set_symm(h, v, step.data_ptr(), (int)rank, (int)n_max);
}
void dstep_add_gemv_b(int64_t h, const at::Tensor& w, const at::Tensor& x, const c10::optional<at::Tensor>& bias,
                      const c10::optional<at::Tensor>& norm_w, double eps, double offset, int64_t act,
                      const c10::optional<at::Tensor>& residual, at::Tensor& y, bool allreduce) {
  TORCH_CHECK(w.is_cuda() && w.dim() == 2 && w.is_contiguous() && is_bf16(w) && is_bf16(x) && is_bf16(y));
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.size(1) == w.size(1) && y.dim() == 2 && y.stride(1) == 1 && x.stride(0) % 8 == 0);
  const int N = w.size(0), K = w.size(1);
  TORCH_CHECK(K % 64 == 0 && y.size(1) == (act != 0 ? N / 2 : N) && y.size(0) == x.size(0));
  if (residual.has_value()) TORCH_CHECK(residual->stride(0) == y.stride(0) && residual->stride(1) == 1 && is_bf16(*residual));
  dstep_add_gemv(h, w.data_ptr(), N, K, x.data_ptr(), (int)x.stride(0), optr(bias), optr(norm_w), (float)eps, (float)offset, (int)act,
                 optr(residual), y.data_ptr(), (int)y.stride(0), allreduce);
}
void dstep_add_attn_b(int64_t h, const at::Tensor& qkv, at::Tensor& out, at::Tensor& k_cache, at::Tensor& v_cache,
                      const c10::optional<at::Tensor>& q_norm, const c10::optional<at::Tensor>& k_norm, double eps, int64_t B,
                      int64_t T, int64_t nq, int64_t nkv, int64_t D, double scale, int64_t window, const c10::optional<at::Tensor>& sinks,
                      int64_t s_hint) {
  TORCH_CHECK(is_bf16(qkv) && is_bf16(out) && is_bf16(k_cache) && k_cache.dim() == 4 && k_cache.is_contiguous() && v_cache.is_contiguous());
  TORCH_CHECK(k_cache.size(1) == nkv && k_cache.size(3) == D && qkv.is_contiguous() && out.is_contiguous());
  const int S = k_cache.size(2), L = k_cache.size(0);
  const int nsplit = pick_nsplit((int)B, (int)nkv, s_hint > 0 ? (int)std::min<int64_t>(s_hint, S) : S);
  dstep_add_attn(h, qkv.data_ptr(), out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), optr(q_norm), optr(k_norm), (float)eps,
                 (int)B, (int)T, (int)nq, (int)nkv, (int)D, S, L, (float)scale, (int)window,
                 sinks.has_value() ? sinks->data_ptr<float>() : nullptr, nsplit);
}
void dstep_launch_b(int64_t h, const at::Tensor& positions, const at::Tensor& write_pos, const at::Tensor& lines,
                    const at::Tensor& cos, const at::Tensor& sin, int64_t call_base, int64_t parity_base) {
  TORCH_CHECK(positions.is_cuda() && positions.scalar_type() == at::kInt && write_pos.scalar_type() == at::kInt &&
              lines.scalar_type() == at::kInt && positions.is_contiguous() && write_pos.is_contiguous() && lines.is_contiguous());
  TORCH_CHECK(cos.scalar_type() == at::kFloat && sin.scalar_type() == at::kFloat && cos.is_contiguous() && sin.is_contiguous());
  c10::cuda::CUDAGuard guard(positions.device());
  dstep_launch(h, positions.data_ptr<int>(), write_pos.data_ptr<int>(), lines.data_ptr<int>(), cos.data_ptr<float>(),
               sin.data_ptr<float>(), (int)call_base, (int)parity_base, cur_stream());
}
}  // namespace nxdi

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("symm_heap_create", [](int64_t bytes, int64_t dev, int64_t world, int64_t rank) { return (int64_t)nxdi::symm_heap_create(bytes, (int)dev, (int)world, (int)rank); });
  m.def("symm_heap_size", [](int64_t h) { return (int64_t)nxdi::symm_heap_size(h); });
  m.def("symm_heap_local_va", [](int64_t h) { return (int64_t)nxdi::symm_heap_local_va(h); });
  m.def("symm_heap_export_fd", [](int64_t h) { return (int64_t)nxdi::symm_heap_export_fd(h); });
  m.def("symm_heap_import_peer", [](int64_t h, int64_t peer, int64_t fd) { return (int64_t)nxdi::symm_heap_import_peer(h, (int)peer, (int)fd); });
  m.def("symm_heap_multicast_supported", [](int64_t dev) { return nxdi::symm_heap_multicast_supported((int)dev); });
  m.def("symm_heap_mc_create", [](int64_t h, int64_t world) { return (int64_t)nxdi::symm_heap_mc_create(h, (int)world); });
  m.def("symm_heap_mc_import", [](int64_t h, int64_t fd) { nxdi::symm_heap_mc_import(h, (int)fd); });
  m.def("symm_heap_mc_add_device", [](int64_t h) { nxdi::symm_heap_mc_add_device(h); });
  m.def("symm_heap_mc_bind_map", [](int64_t h) { return (int64_t)nxdi::symm_heap_mc_bind_map(h); });
  m.def("symm_heap_destroy", [](int64_t h) { nxdi::symm_heap_destroy(h); });
  m.def("nvls_collective", &nxdi::nvls_collective);
  m.def("attention_prefill_tc", &nxdi::attention_prefill_tc);
  m.def("gemm_reduce_scatter", &nxdi::gemm_reduce_scatter);
  m.def("rmsnorm_quant", &nxdi::rmsnorm_quant);
  m.def("gemm_fp8", &nxdi::gemm_fp8);
  m.def("dstep_new", &nxdi::dstep_new_b);
  m.def("dstep_set_symm", &nxdi::dstep_set_symm_b);
  m.def("dstep_add_gemv", &nxdi::dstep_add_gemv_b);
  m.def("dstep_add_attn", &nxdi::dstep_add_attn_b);
  m.def("dstep_finalize", [](int64_t h) { nxdi::dstep_finalize(h); });
  m.def("dstep_num_allreduce", [](int64_t h) { return (int64_t)nxdi::dstep_num_allreduce(h); });
  m.def("dstep_free", [](int64_t h) { nxdi::dstep_free(h); });
  m.def("dstep_launch", &nxdi::dstep_launch_b);
  m.def("rmsnorm", &nxdi::rmsnorm);
  m.def("gemv", &nxdi::gemv);
  m.def("gemv_allreduce", &nxdi::gemv_allreduce);
  m.def("set_prof", [](const c10::optional<at::Tensor>& buf) {
    if (!buf.has_value()) { nxdi::gemv2_set_prof(nullptr, 0); return; }
    TORCH_CHECK(buf->is_cuda() && buf->scalar_type() == at::kLong && buf->is_contiguous());
    nxdi::gemv2_set_prof(reinterpret_cast<unsigned long long*>(buf->data_ptr<int64_t>()), buf->numel() / (148 * 8));
  });
  m.def("prof_count", []() { return (int64_t)nxdi::prof_count(); });
  m.def("gemv2_supported", [](int64_t T, int64_t K, int64_t wt) { return nxdi::gemv2_supported((int)T, (int)K, (int)wt); },
        pybind11::arg("T"), pybind11::arg("K"), pybind11::arg("wt") = 0);
  m.def("moe_decode", &nxdi::moe_decode);
  m.def("moe_grouped", &nxdi::moe_grouped);
  m.def("dequant_bf16", &nxdi::dequant_bf16);
  m.def("gemm", &nxdi::gemm);
  m.def("symm_alloc", &nxdi::symm_alloc);
  m.def("symm_open", &nxdi::symm_open);
  m.def("symm_close", &nxdi::symm_close);
  m.def("symm_free", &nxdi::symm_free);
  m.def("symm_as_tensor", &nxdi::symm_as_tensor);
  m.def("rope_kv_append", &nxdi::rope_kv_append);
  m.def("rope_kv_split_append", &nxdi::rope_kv_split_append);
  m.def("kv_append", &nxdi::kv_append);
  m.def("paged_kv_append", &nxdi::paged_kv_append);
  m.def("argmax", &nxdi::argmax);
  m.def("topk_sample", &nxdi::topk_sample);
  m.def("attention_decode", &nxdi::attention_decode);
  m.def("paged_attention_decode", &nxdi::paged_attention_decode);
  m.def("rope_attention_decode", &nxdi::rope_attention_decode);
  m.def("attention_prefill", &nxdi::attention_prefill);
}
