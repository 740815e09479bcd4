"""Numerics of every sm_100a kernel against the plain-PyTorch fp32 definition of the same op."""
import math

import pytest
import torch

from neuronx_distributed_inference_b200 import ops
from neuronx_distributed_inference_b200.ops import reference as ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-9)).item()


@pytest.fixture(scope="module", autouse=True)
def _ext():
    from neuronx_distributed_inference_b200.ops._ext import load_extension
    load_extension()


@pytest.mark.parametrize("rows,H", [(1, 4096), (7, 4096), (33, 8192), (4, 1024)])
def test_rmsnorm(rows, H):
    x = torch.randn(rows, H, device=DEV, dtype=torch.bfloat16)
    w = (torch.randn(H, device=DEV) * 0.1 + 1).to(torch.bfloat16)
    y = ops.rmsnorm(x, w, 1e-5)
    yr = ref.rmsnorm(x.float(), w.float(), 1e-5)
    assert _rel(y, yr) < 5e-3
    r = torch.randn_like(x)
    y2, r2 = ops.rmsnorm(x, w, 1e-5, 0.0, r)
    yr2, rr2 = ref.rmsnorm(x, w, 1e-5, 0.0, r)
    assert _rel(r2, rr2) < 1e-6 and _rel(y2, yr2) < 8e-3


@pytest.mark.parametrize("T", [1, 2, 3, 8])
@pytest.mark.parametrize("N,K", [(4096, 4096), (6144, 4096), (1000 * 16, 2048), (4096, 14336), (512, 512), (40, 256),
                                 (3072, 4096), (64128, 4096), (4096, 7168), (4096, 2048), (4096, 1792), (768, 4096), (4096, 320), (400, 512)])
def test_gemv(T, N, K):
    x = torch.randn(T, K, device=DEV, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    y = ops.linear(x, w, b)
    yr = ref.linear(x.float(), w.float(), b.float())
    assert y.shape == yr.shape and _rel(y, yr) < 6e-3


@pytest.mark.parametrize("T", [1, 2, 5])
@pytest.mark.parametrize("N,K", [(2 * 14336, 4096), (2 * 1792, 4096), (2 * 11008, 4096), (14336, 4096)])
def test_gemv_norm_swiglu(T, N, K):
    x = torch.randn(T, K, device=DEV, dtype=torch.bfloat16) * 3
    g = (torch.randn(K, device=DEV) * 0.2 + 1).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    y = ops.linear(x, w, None, norm_weight=g, norm_eps=1e-5, act="silu_mul")
    xn = ref.rmsnorm(x.float(), g.float(), 1e-5).to(torch.bfloat16).float()
    yr = ref.linear(xn, w.float(), None, act="silu_mul")
    assert y.shape == (T, N // 2) and _rel(y, yr) < 8e-3


def test_gemv_x_global_path():
    # T*K too large for shared memory -> pre-normalised x read through L1
    T, N, K = 8, 1024, 14336
    x = torch.randn(T, K, device=DEV, dtype=torch.bfloat16)
    g = torch.ones(K, device=DEV, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    y = ops.linear(x, w, None, norm_weight=g, norm_eps=1e-5)
    yr = ref.linear(ref.rmsnorm(x.float(), g.float(), 1e-5).to(torch.bfloat16).float(), w.float())
    assert _rel(y, yr) < 8e-3


@pytest.mark.parametrize("D,nq,nkv", [(128, 32, 8), (64, 8, 8), (128, 4, 1)])
@pytest.mark.parametrize("T", [1, 3])
def test_rope_kv_append(D, nq, nkv, T):
    B, S, L = 3, 64, 4
    qkv = torch.randn(B, T, (nq + 2 * nkv) * D, device=DEV, dtype=torch.bfloat16)
    pos = torch.randint(0, S - T, (B, 1), device=DEV) + torch.arange(T, device=DEV)
    pos[1, -1] = -1  # skipped write
    ang = torch.rand(B, T, D // 2, device=DEV) * 6
    cos, sin = ang.cos(), ang.sin()
    lines = torch.tensor([2, 0, 7], device=DEV, dtype=torch.int32)  # 7 is out of range -> skipped
    kc = torch.zeros(L, nkv, S, D, device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    qn = (torch.randn(D, device=DEV) * 0.1 + 1).to(torch.bfloat16)
    kn = (torch.randn(D, device=DEV) * 0.1 + 1).to(torch.bfloat16)
    for norms in (False, True):
        kc.zero_(); vc.zero_()
        q = ops.rope_kv_append(qkv, cos, sin, kc, vc, lines, pos.int(), nq, nkv, D, False,
                               qn if norms else None, kn if norms else None, 1e-6)
        kr, vr = torch.zeros_like(kc), torch.zeros_like(vc)
        ops.set_kernels_enabled(False)
        qr = ops.rope_kv_append(qkv, cos, sin, kr, vr, lines, pos.int(), nq, nkv, D, False,
                                qn if norms else None, kn if norms else None, 1e-6)
        ops.set_kernels_enabled(True)
        assert _rel(q, qr) < 1e-2 and _rel(kc, kr) < 1e-2 and torch.equal(vc, vr)


def test_kv_append_and_paged():
    B, T, H, D, S, L = 2, 5, 4, 128, 32, 3
    k = torch.randn(B, T, H, D, device=DEV, dtype=torch.bfloat16)
    v = torch.randn_like(k)
    pos = torch.tensor([[3, 4, 5, 6, -1], [0, 1, 2, 31, 40]], device=DEV, dtype=torch.int32)
    lines = torch.tensor([1, 2], device=DEV, dtype=torch.int32)
    kc = torch.zeros(L, H, S, D, device=DEV, dtype=torch.bfloat16); vc = torch.zeros_like(kc)
    kr = torch.zeros_like(kc); vr = torch.zeros_like(kc)
    ops.kv_append(kc, vc, k, v, lines, pos)
    ref.kv_append(kr, vr, k, v, lines, pos)
    assert torch.equal(kc, kr) and torch.equal(vc, vr)
    nb, bs = 6, 8
    pk = torch.zeros(nb, bs, H, D, device=DEV, dtype=torch.bfloat16); pv = torch.zeros_like(pk)
    rk = torch.zeros_like(pk); rv = torch.zeros_like(pk)
    slots = torch.tensor([[0, 1, 9, 17, -1], [47, 5, 6, 7, 8]], device=DEV, dtype=torch.int32)
    ops.paged_kv_append(pk, pv, k, v, slots)
    ref.paged_kv_append(rk, rv, k, v, slots)
    assert torch.equal(pk, rk) and torch.equal(pv, rv)


@pytest.mark.parametrize("D", [128, 64])
@pytest.mark.parametrize("Hq,Hkv,T", [(32, 8, 1), (4, 1, 1), (8, 8, 2), (16, 2, 5), (32, 8, 8)])
@pytest.mark.parametrize("window", [0, 37])
def test_attention_decode(D, Hq, Hkv, T, window):
    if T * (Hq // Hkv) > 64:
        pytest.skip("rows > 64")
    B, S, L = 3, 700, 4
    kc = torch.randn(L, Hkv, S, D, device=DEV, dtype=torch.bfloat16)
    vc = torch.randn_like(kc)
    q = torch.randn(B, T, Hq, D, device=DEV, dtype=torch.bfloat16)
    base = torch.tensor([[5], [300], [S - T - 1]], device=DEV)
    pos = (base + torch.arange(T, device=DEV)).int()
    lines = torch.tensor([3, 1, 0], device=DEV, dtype=torch.int32)
    scale = 1 / math.sqrt(D)
    o = ops.attention_decode(q, kc, vc, lines, pos, scale, window or None)
    orf = ref.attention_decode(q.float(), kc.float(), vc.float(), lines, pos, scale, window or None)
    assert _rel(o, orf) < 1.5e-2


def test_attention_decode_sinks_and_masked_line():
    B, T, Hq, Hkv, D, S, L = 2, 1, 8, 2, 64, 256, 2
    kc = torch.randn(L, Hkv, S, D, device=DEV, dtype=torch.bfloat16); vc = torch.randn_like(kc)
    q = torch.randn(B, T, Hq, D, device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([[100], [7]], device=DEV, dtype=torch.int32)
    lines = torch.tensor([1, 0], device=DEV, dtype=torch.int32)
    sinks = torch.randn(Hq, device=DEV)
    o = ops.attention_decode(q, kc, vc, lines, pos, 0.125, None, None, sinks)
    orf = ref.attention_decode(q.float(), kc.float(), vc.float(), lines, pos, 0.125, None, None, sinks)
    assert _rel(o, orf) < 1.5e-2


@pytest.mark.parametrize("D", [128, 64])
@pytest.mark.parametrize("B,T,Hq,Hkv", [(2, 128, 8, 2), (1, 200, 4, 4), (2, 33, 8, 1), (1, 1024, 4, 1)])
@pytest.mark.parametrize("window", [0, 50])
def test_attention_prefill(D, B, T, Hq, Hkv, window):
    q = torch.randn(B, T, Hq, D, device=DEV, dtype=torch.bfloat16)
    k = torch.randn(B, T, Hkv, D, device=DEV, dtype=torch.bfloat16)
    v = torch.randn_like(k)
    scale = 1 / math.sqrt(D)
    o = ops.attention_prefill(q, k, v, scale, True, window or None)
    orf = ref.attention_prefill(q.float(), k.float(), v.float(), scale, True, window or None)
    assert _rel(o, orf) < 1.5e-2


def test_paged_attention_decode():
    B, T, Hq, Hkv, D, bs, nb = 2, 2, 8, 2, 128, 16, 40
    kc = torch.randn(nb, bs, Hkv, D, device=DEV, dtype=torch.bfloat16); vc = torch.randn_like(kc)
    q = torch.randn(B, T, Hq, D, device=DEV, dtype=torch.bfloat16)
    bt = torch.randperm(nb, device=DEV)[:B * 10].view(B, 10).int()
    pos = torch.tensor([[100, 101], [17, 18]], device=DEV, dtype=torch.int32)
    o = ops.paged_attention_decode(q, kc, vc, bt, pos, 0.09)
    orf = ref.paged_attention_decode(q.float(), kc.float(), vc.float(), bt, pos, 0.09)
    assert _rel(o, orf) < 1.5e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_argmax(dtype):
    x = torch.randn(5, 128256, device=DEV).to(dtype)
    assert torch.equal(ops.argmax(x), x.float().argmax(-1))
    x = torch.randn(3, 1000, device=DEV).to(dtype)
    assert torch.equal(ops.argmax(x), x.float().argmax(-1))


def test_topk_sample_matches_reference():
    B, V = 6, 32000
    logits = torch.randn(B, V, device=DEV) * 3
    top_k = torch.tensor([1, 5, 50, 0, 256, 10], device=DEV, dtype=torch.int32)
    top_p = torch.tensor([1.0, 0.9, 0.5, 1.0, 0.95, 1.0], device=DEV)
    temp = torch.tensor([1.0, 0.7, 1.3, 1.0, 0.0, 2.0], device=DEV)
    for seed in range(4):
        torch.manual_seed(seed)
        rand = torch.rand(B, device=DEV)
        got = ops.sample(logits, top_k, top_p, temp, rand, 256)
        exp = ref.sample(logits, top_k, top_p, temp, rand, 256)
        assert torch.equal(got, exp), (got, exp)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 6144, 4096), (256, 4096, 14336), (77, 1000 * 8, 512), (1024, 4096, 4096),
                                   (9, 256, 256)])
def test_gemm_tcgen05(M, N, K):
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=DEV, dtype=torch.bfloat16)
    y = ops.linear(x, w, b, residual=r)
    yr = ref.linear(x.float(), w.float(), b.float()) + r.float()
    assert ops.stats["gemm_tcgen05"] > 0
    assert y.shape == yr.shape and _rel(y, yr) < 6e-3


@pytest.mark.parametrize("M,N,K", [(256, 2 * 14336, 4096), (100, 2 * 1792, 1024)])
def test_gemm_tcgen05_norm_swiglu(M, N, K):
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16) * 2
    g = (torch.randn(K, device=DEV) * 0.2 + 1).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    y = ops.linear(x, w, None, norm_weight=g, norm_eps=1e-5, act="silu_mul")
    xn = ref.rmsnorm(x.float(), g.float(), 1e-5).to(torch.bfloat16).float()
    yr = ref.linear(xn, w.float(), None, act="silu_mul")
    assert y.shape == (M, N // 2) and _rel(y, yr) < 8e-3


@pytest.mark.parametrize("D,nq,nkv,T,ctx,qk_norm", [(128, 32, 8, 1, 200, False), (128, 8, 2, 4, 1500, False), (64, 14, 2, 1, 77, True),
                                                     (128, 4, 1, 8, 40, True),
                                                     # multi-head attention (one q head per kv head: a single row per CTA), short contexts
                                                     (128, 4, 4, 1, 12, False), (64, 8, 8, 1, 50, False), (128, 32, 32, 1, 130, False)])
def test_rope_attention_decode_equals_two_kernel_path(D, nq, nkv, T, ctx, qk_norm):
    """RoPE + q/k norm + cache append folded into the attention kernel == rope_kv_append kernel + attention kernel."""
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    B, S, L = 3, 2048, 4
    qkv = torch.randn(B, T, (nq + 2 * nkv) * D, device=dev, dtype=dt)
    kc = torch.randn(L, nkv, S, D, device=dev, dtype=dt)
    vc = torch.randn(L, nkv, S, D, device=dev, dtype=dt)
    lines = torch.tensor([2, 0, 3], device=dev, dtype=torch.int32)
    base = torch.tensor([ctx, ctx // 2, 5], device=dev, dtype=torch.int32).view(B, 1)
    pos = base + torch.arange(T, device=dev, dtype=torch.int32).view(1, T)
    ang = torch.rand(B, T, D // 2, device=dev) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    qn = (1 + 0.1 * torch.randn(D, device=dev)).to(dt) if qk_norm else None
    kn = (1 + 0.1 * torch.randn(D, device=dev)).to(dt) if qk_norm else None
    k1, v1, k2, v2 = kc.clone(), vc.clone(), kc.clone(), vc.clone()
    q = ops.rope_kv_append(qkv, cos, sin, k1, v1, lines, pos, nq, nkv, D, False, qn, kn, 1e-6)
    exp = ops.attention_decode(q, k1, v1, lines, pos, D ** -0.5, None, None, None, seq_hint=S)
    got = ops.rope_attention_decode(qkv, cos, sin, k2, v2, lines, pos, pos, nq, nkv, D, D ** -0.5, None, None, qn, kn, 1e-6, seq_hint=S)
    assert ops.stats["rope_attn_decode"] > 0
    assert torch.equal(k1, k2) and torch.equal(v1, v2)
    assert (got.float() - exp.float()).abs().max().item() <= 2e-2
    # and against the fp32 PyTorch definition of the whole step (RoPE, append, attention over the cache lines)
    qr, kr, vr = qkv.view(B, T, nq + 2 * nkv, D).float().split([nq, nkv, nkv], 2)
    if qk_norm:
        qr, kr = ref.rmsnorm(qr, qn.float(), 1e-6), ref.rmsnorm(kr, kn.float(), 1e-6)
    qr, kr = ref.apply_rope(qr, cos, sin, False), ref.apply_rope(kr, cos, sin, False)
    k3, v3 = kc.float().clone(), vc.float().clone()
    ref.kv_append(k3, v3, kr, vr, lines, pos)
    full = ref.attention_decode(qr, k3, v3, lines, pos, D ** -0.5)
    assert _rel(got, full) < 2e-2, _rel(got, full)


@pytest.mark.parametrize("D,nq,nkv,B,T,qk_norm", [(128, 8, 2, 2, 300, False), (64, 6, 2, 1, 77, True), (128, 4, 4, 3, 16, True)])
def test_prefill_split_rope_append_single_kernel(D, nq, nkv, B, T, qk_norm):
    """csrc/rope_kv.cu prefill flavour: q, rotated k, v (contiguous) + cache write from ONE pass over the fused QKV output."""
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    L, S = B + 1, 512
    qkv = torch.randn(B, T, (nq + 2 * nkv) * D, device=dev, dtype=dt)
    kc = torch.zeros(L, nkv, S, D, device=dev, dtype=dt)
    vc = torch.zeros_like(kc)
    lines = torch.arange(B, device=dev, dtype=torch.int32)
    pos = torch.arange(T, device=dev, dtype=torch.int32).view(1, T).repeat(B, 1)
    pos[0, -3:] = -1                                      # padding: not written to the cache
    ang = torch.rand(B, T, D // 2, device=dev) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    qn = (1 + 0.1 * torch.randn(D, device=dev)).to(dt) if qk_norm else None
    kn = (1 + 0.1 * torch.randn(D, device=dev)).to(dt) if qk_norm else None
    q, k, v = ops.rope_kv_split_append(qkv, cos, sin, kc, vc, lines, pos, nq, nkv, D, qn, kn, 1e-6)
    k2, v2 = torch.zeros_like(kc), torch.zeros_like(vc)
    q_ref = ops.rope_kv_append(qkv, cos, sin, k2, v2, lines, pos, nq, nkv, D, False, qn, kn, 1e-6)     # the decode-time kernel
    assert torch.equal(q, q_ref) and torch.equal(kc, k2) and torch.equal(vc, v2)
    qr, kr, vr = qkv.view(B, T, nq + 2 * nkv, D).split([nq, nkv, nkv], 2)
    if qk_norm:
        kr = ref.rmsnorm(kr, kn, 1e-6)
    kr = ref.apply_rope(kr, cos, sin, False)
    assert (k.float() - kr.float()).abs().max().item() < 3e-2 and torch.equal(v, vr)
    assert torch.equal(kc[B - 1, :, :T - 3], k[B - 1, :T - 3].transpose(0, 1)) and float(kc[0, :, T - 3:T].abs().max()) == 0.0


@pytest.mark.parametrize("T,k,E,H,I,off", [(2, 2, 8, 4096, 1792, 0), (1, 4, 16, 2048, 768, 0), (8, 2, 4, 1024, 512, 2), (3, 8, 64, 2048, 256, 16)])
def test_moe_decode_kernels_match_reference(T, k, E, H, I, off):
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    x = torch.randn(T, H, device=dev, dtype=dt)
    wgu = (torch.randn(E, 2 * I, H, device=dev) * 0.03).to(dt)
    wd = (torch.randn(E, H, I, device=dev) * 0.03).to(dt)
    n_global = E + 2 * off                     # some slots route to experts owned by other ranks
    idx = torch.stack([torch.randperm(n_global, device=dev)[:k] for _ in range(T)])
    w = torch.rand(T, k, device=dev)
    before = ops.stats["moe_decode"]
    got = ops.moe_experts(x, wgu, wd, w, idx, "silu_mul", off)
    assert ops.stats["moe_decode"] == before + 1
    exp = ref.moe_experts(x, wgu, wd, w, idx, "silu_mul", off)
    err = (got.float() - exp.float()).abs().max().item()
    assert err <= 0.02 * exp.float().abs().max().item() + 0.02, err


@pytest.mark.parametrize("N,k,E,H,I,off,act,bias,scale_input", [
    (300, 2, 8, 1024, 512, 0, "silu_mul", False, False),      # every expert a few partial tiles
    (64, 4, 16, 2048, 768, 0, "silu_mul", False, False),      # fewer rows than experts x 128: mostly padding
    (1000, 2, 4, 1024, 448, 2, "gelu_tanh_mul", True, False), # expert-parallel shard (foreign experts), biases, odd I (7 x 64)
    (257, 1, 16, 512, 256, 0, "silu_mul", False, True),       # Llama-4 style: top-1, affinity on the expert input
    (2048, 4, 16, 2048, 1344, 0, "silu_mul", False, False),   # DBRX TP8 shard shape (I = 10752 / 8) at reduced hidden size
    (200, 4, 8, 512, 256, 0, "gpt_oss_glu", True, False),     # GPT-OSS: biased experts, clamped SwiGLU epilogue (callable with a kernel twin)
    (3, 4, 8, 512, 256, 0, "gpt_oss_glu", True, False),       # ... at decode size (no moe_decode kernel for biased experts: grouped path)
])
def test_moe_grouped_gemm_matches_reference(N, k, E, H, I, off, act, bias, scale_input):
    """Prefill MoE: device-side permutation + grouped tcgen05 GEMMs (csrc/moe_grouped.cu, gemm_tcgen05.cu grouped mode)."""
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    x = torch.randn(N, H, device=dev, dtype=dt)
    wgu = (torch.randn(E, 2 * I, H, device=dev) * 0.03).to(dt)
    wd = (torch.randn(E, H, I, device=dev) * 0.03).to(dt)
    gb = (torch.randn(E, 2 * I, device=dev) * 0.1).to(dt) if bias else None
    db = (torch.randn(E, H, device=dev) * 0.1).to(dt) if bias else None
    n_global = E + 2 * off
    idx = torch.rand(N, n_global, device=dev).topk(k, dim=-1).indices
    if N >= 1000:
        idx[:, 0] = off          # a hot expert: many tiles of one expert, others nearly empty
    w = torch.rand(N, k, device=dev)
    act_fn = None
    if act == "gpt_oss_glu":
        from neuronx_distributed_inference_b200.models.gpt_oss.modeling_gpt_oss import gpt_oss_glu
        act, act_fn = "silu_mul", gpt_oss_glu
        x = x * 4                                # reach the clamps
    before = ops.stats["moe_grouped"]
    got = ops.moe_experts(x, wgu, wd, w, idx, act, off, gb, db, act_fn, scale_input)
    assert ops.stats["moe_grouped"] == before + 1
    got2 = ops.moe_experts(x, wgu, wd, w, idx, act, off, gb, db, act_fn, scale_input)
    assert torch.equal(got, got2)                # row order inside an expert may differ run to run; the result may not
    exp = ref.moe_experts(x, wgu, wd, w, idx, act, off, gb, db, act_fn, scale_input)
    err = (got.float() - exp.float()).abs().max().item()
    assert err <= 0.02 * exp.float().abs().max().item() + 0.02, err
    # the same launch sequence replays under a CUDA graph with a different routing
    idx_s, w_s, x_s = idx.clone(), w.clone(), x.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.moe_experts(x_s, wgu, wd, w_s, idx_s, act, off, gb, db, act_fn, scale_input)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_s = ops.moe_experts(x_s, wgu, wd, w_s, idx_s, act, off, gb, db, act_fn, scale_input)
    idx2 = torch.rand(N, n_global, device=dev).topk(k, dim=-1).indices
    idx_s.copy_(idx2)
    g.replay()
    exp2 = ref.moe_experts(x, wgu, wd, w, idx2, act, off, gb, db, act_fn, scale_input)
    err2 = (out_s.float() - exp2.float()).abs().max().item()
    assert err2 <= 0.02 * exp2.float().abs().max().item() + 0.02, err2


def test_moe_grouped_fp8_experts_w8a8():
    """fp8 expert banks + dynamic per-row activation quantisation through the grouped kind::f8f6f4 GEMMs == the same math in fp32."""
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    N, k, E, H, I = 600, 2, 8, 1024, 512
    x = torch.randn(N, H, device=dev, dtype=dt)
    wgu = torch.randn(E, 2 * I, H, device=dev) * 0.03
    wd = torch.randn(E, H, I, device=dev) * 0.03
    gs = wgu.abs().amax(-1) / 448.0
    ds = wd.abs().amax(-1) / 448.0
    wgu_q = (wgu / gs.unsqueeze(-1)).to(torch.float8_e4m3fn)
    wd_q = (wd / ds.unsqueeze(-1)).to(torch.float8_e4m3fn)
    idx = torch.rand(N, E, device=dev).topk(k, dim=-1).indices
    w = torch.rand(N, k, device=dev)
    ops.set_activation_quant(True)
    try:
        before = ops.stats["moe_grouped_fp8"]
        got = ops.moe_experts(x, wgu_q, wd_q, w, idx, "silu_mul", 0, None, None, None, False, gs, ds)
        assert ops.stats["moe_grouped_fp8"] == before + 1
    finally:
        ops.set_activation_quant(False)
    # reference: quantise the activations the same way (per row, e4m3), accumulate in fp32
    def q8(t):
        s_ = t.float().abs().amax(-1, keepdim=True).clamp_min(1e-12) / 448.0
        return (t.float() / s_).to(torch.float8_e4m3fn).float() * s_
    out = torch.zeros(N, H, device=dev)
    for e in range(E):
        sel = idx == e
        tok = sel.any(-1).nonzero().flatten()
        if tok.numel() == 0:
            continue
        wt = (w * sel).sum(-1)[tok]
        hh = q8(x[tok]) @ (wgu_q[e].float() * gs[e].unsqueeze(-1)).t()
        g_, u_ = hh.chunk(2, -1)
        hh = (torch.nn.functional.silu(g_) * u_).to(dt)
        yy = (q8(hh) @ (wd_q[e].float() * ds[e].unsqueeze(-1)).t()).to(dt).float()
        out[tok] += yy * wt.unsqueeze(-1)
    err = (got.float() - out).abs().max().item()
    assert err <= 0.03 * out.abs().max().item() + 0.03, err


@pytest.mark.parametrize("wdtype", [torch.int8, torch.float8_e4m3fn])
@pytest.mark.parametrize("T,N,K,act,norm,res,per_tensor", [(2, 4096, 4096, None, True, False, False), (1, 6144, 4096, None, False, True, False),
                                                           (8, 2048, 14336, None, False, True, True), (4, 7168, 4096, "silu_mul", True, False, False),
                                                           (3, 1000, 1040, None, False, False, False),
                                                           # Llama-2-7b: K = 11008 is not a multiple of the 1024-k stage, GLU tiles
                                                           (2, 4096, 11008, None, False, True, False), (2, 22016, 4096, "silu_mul", True, False, False),
                                                           (5, 520, 1280, None, True, True, False)])
def test_quantized_gemv_matches_dequantized_reference(wdtype, T, N, K, act, norm, res, per_tensor):
    """Weight-only int8 / fp8 decode GEMV == dequantise-then-matmul.  K % 128 == 0: the TMA-streamed gemv2 kernel (8-bit weights
    expanded to f16 in registers, csrc/gemv2_body.cuh); otherwise the CUDA-core fallback (csrc/qgemv.cu)."""
    torch.manual_seed(0)
    dev, dt = "cuda", torch.bfloat16
    w = torch.randn(N, K, device=dev) * 0.05
    if per_tensor:
        q, s = ref.quantize_per_tensor(w, wdtype)
        s = s.reshape(1).float()
    else:
        q, s = ref.quantize_per_channel(w, wdtype)
    x = torch.randn(T, K, device=dev, dtype=dt)
    nw = (1 + 0.1 * torch.randn(K, device=dev)).to(dt) if norm else None
    n_out = N // 2 if act else N
    r = torch.randn(T, n_out, device=dev, dtype=dt) if res else None
    bias = (torch.randn(N, device=dev) * 0.1).to(dt)
    before = ops.stats["qgemv"]
    got = ops.linear(x, q, bias, norm_weight=nw, norm_eps=1e-5, act=act, scale=s, residual=r)
    assert ops.stats["qgemv"] == before + 1
    exp = ref.linear(x.float(), q, bias.float(), nw.float() if norm else None, 1e-5, 0.0, act, s)
    if r is not None:
        exp = exp + r.float()
    err = (got.float() - exp.float()).abs().max().item()
    assert err <= 0.02 * exp.abs().max().item() + 0.03, err


@pytest.mark.parametrize("wdtype", [torch.int8, torch.float8_e4m3fn])
def test_weight_only_prefill_expands_once_then_tensor_core_gemm(wdtype):
    """T > 8 with 8-bit weights and no activation quantisation: one dequantise-to-bf16 pass into the shared scratch + tcgen05 GEMM."""
    torch.manual_seed(0)
    T, N, K = 256, 2 * 1408, 1024
    q, s = ref.quantize_per_channel(torch.randn(N, K, device="cuda") * 0.05, wdtype)
    x = torch.randn(T, K, device="cuda", dtype=torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    n0, n1 = ops.stats["dequant_gemm"], ops.stats["gemm_tcgen05"]
    got = ops.linear(x, q, None, norm_weight=nw, norm_eps=1e-5, act="silu_mul", scale=s)
    assert ops.stats["dequant_gemm"] == n0 + 1 and ops.stats["gemm_tcgen05"] == n1 + 1
    exp = ref.linear(x.float(), q, None, nw.float(), 1e-5, 0.0, "silu_mul", s)
    assert _rel(got, exp) < 1e-2, _rel(got, exp)


@pytest.mark.parametrize("B,T,Hq,Hkv,D", [(2, 200, 8, 8, 64), (1, 1500, 4, 4, 128), (3, 77, 16, 4, 128)])
def test_bidirectional_flash_attention_matches_reference(B, T, Hq, Hkv, D):
    """Encoder (non-causal) mode of the flash kernel: vision towers, Whisper encoder, FLUX joint attention."""
    torch.manual_seed(0)
    q = torch.randn(B, T, Hq, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, T, Hkv, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, T, Hkv, D, device="cuda", dtype=torch.bfloat16)
    stat = "attn_prefill_tc" if D == 128 else "attn_prefill"      # head_dim 128: the tcgen05 kernel, 64: the mma.sync one
    before = ops.stats[stat]
    got = ops.attention_prefill(q, k, v, D ** -0.5, causal=False)
    assert ops.stats[stat] == before + 1
    exp = ref.attention_prefill(q.float(), k.float(), v.float(), D ** -0.5, causal=False)
    assert (got.float() - exp).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,T,Hq,Hkv,causal,window,softcap,sinks", [
    (1, 128, 2, 1, True, 0, 0.0, False), (2, 300, 4, 2, True, 0, 0.0, False), (1, 1024, 8, 2, True, 0, 0.0, False),
    (1, 640, 4, 4, True, 200, 0.0, True), (2, 257, 2, 1, False, 0, 0.0, False), (1, 512, 4, 1, True, 0, 30.0, False),
    (1, 2048, 4, 1, True, 0, 0.0, False)])
def test_attention_prefill_tcgen05(B, T, Hq, Hkv, causal, window, softcap, sinks):
    """csrc/attention_tc.cu (tcgen05 QK^T / PV, TMEM accumulators, TMA tiles, MN-major V) vs the fp32 PyTorch definition."""
    D = 128
    torch.manual_seed(1)
    q = torch.randn(B, T, Hq, D, device=DEV, dtype=torch.bfloat16)
    k = torch.randn(B, T, Hkv, D, device=DEV, dtype=torch.bfloat16)
    v = torch.randn(B, T, Hkv, D, device=DEV, dtype=torch.bfloat16)
    sk = torch.randn(Hq, device=DEV) if sinks else None
    scale = D ** -0.5
    n0 = ops.stats["attn_prefill_tc"]
    o = ops.attention_prefill(q, k, v, scale, causal, window or None, sinks=sk, softcap=softcap or None)
    assert ops.stats["attn_prefill_tc"] == n0 + 1
    orf = ref.attention_prefill(q.float(), k.float(), v.float(), scale, causal, window or None, sinks=sk, softcap=softcap or None)
    assert o.shape == orf.shape and _rel(o, orf) < 1.2e-2, _rel(o, orf)


@pytest.mark.parametrize("M,N,K,act,per_tensor", [(256, 4096, 4096, None, False), (300, 2 * 1024, 2048, "silu_mul", False),
                                                  (2048, 4096, 1024, None, True), (64, 512, 512, None, False)])
def test_fp8_w8a8_gemm_and_rmsnorm_quant(M, N, K, act, per_tensor):
    """RMSNorm + dynamic per-token fp8 quantisation (csrc/quant.cu) and the fp8 tensor-core GEMM (tcgen05 kind::f8f6f4) against
    the SAME quantised operands evaluated in fp32 (isolates the kernels from the quantisation error itself)."""
    torch.manual_seed(0)
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16) * 2
    g = (torch.randn(K, device=DEV) * 0.2 + 1).to(torch.bfloat16)
    wf = torch.randn(N, K, device=DEV) / math.sqrt(K)
    ws = (wf.abs().amax() / 448).reshape(1) if per_tensor else wf.abs().amax(1) / 448
    wq = (wf / (ws if per_tensor else ws[:, None])).to(torch.float8_e4m3fn)
    b = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    xq, a_s = ops.rmsnorm_quant(x, g, 1e-5)
    assert xq.dtype == torch.float8_e4m3fn and a_s.shape == (M, 1)
    xn = ref.rmsnorm(x.float(), g.float(), 1e-5).to(torch.bfloat16).float()
    assert _rel(xq.float() * a_s, xn) < 4e-2                      # fp8 e4m3 rounding of the normalised row
    assert torch.allclose(a_s.view(-1), xn.abs().amax(1) / 448, rtol=2e-2)
    ops.set_activation_quant(True)
    try:
        n0 = ops.stats["gemm_fp8"]
        y = ops.linear(x, wq, b, norm_weight=g, norm_eps=1e-5, act=act, scale=ws.float())
        assert ops.stats["gemm_fp8"] == n0 + 1
    finally:
        ops.set_activation_quant(False)
    yr = ref.linear((xq.float() * a_s), wq.float() * (ws if per_tensor else ws[:, None]), b.float(), act=act)
    assert y.shape == yr.shape and _rel(y, yr) < 8e-3, _rel(y, yr)
