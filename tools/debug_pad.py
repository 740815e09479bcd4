import os, sys, torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops
from neuronx_distributed_inference_b200.utils.testing import build_random_llama

def build(hf, pad):
    os.environ["NXDI_B200_PAD_HEAD_DIM"] = "1" if pad else "0"
    return build_random_llama(hf, batch_size=2, seq_len=64, max_context_length=16, device="cuda", dtype="bfloat16", seed=11, output_logits=True)

def run(app, ids, mask, n=3):
    app.reset()
    out = app(ids, attention_mask=mask)
    lg = [out.logits[:, -1].float()]
    tok = out.tokens.view(2, 1)
    pos = torch.full((2, 1), ids.shape[1], dtype=torch.int32)
    for i in range(n):
        out = app(tok, position_ids=pos + i)
        lg.append(out.logits[:, -1].float())
        tok = out.tokens.view(2, 1)
    return torch.stack(lg)

for head_dim, heads, kv in [(100, 4, 4), (80, 4, 2), (48, 8, 2)]:
    hf = dict(hidden_size=heads * head_dim, intermediate_size=512, num_hidden_layers=2, num_attention_heads=heads,
              num_key_value_heads=kv, head_dim=head_dim, vocab_size=512)
    torch.manual_seed(0)
    ids = torch.randint(1, 512, (2, 12)); mask = torch.ones_like(ids)
    a0 = build(hf, False)
    l0 = run(a0, ids, mask)
    ops.set_kernels_enabled(False)
    l0t = run(a0, ids, mask)
    ops.set_kernels_enabled(True)
    sd = {k: v.clone() for k, v in a0.model.state_dict().items()}
    a1 = build(hf, True)
    with torch.no_grad():
        for name, p_ in a1.model.named_parameters():
            src = sd[name]
            p_.copy_(p_.shard_fn(src.cpu(), 0).to(p_.device) if (src.shape != p_.shape and hasattr(p_, "shard_fn")) else src)
    l1 = run(a1, ids, mask)
    ops.set_kernels_enabled(False)
    l1t = run(a1, ids, mask)
    ops.set_kernels_enabled(True)
    rel = lambda a, b: [round(((a[i] - b[i]).norm() / b[i].norm()).item(), 4) for i in range(a.shape[0])]
    print(f"D={head_dim}: unpadded kernels-vs-torch {rel(l0, l0t)}")
    print(f"D={head_dim}: padded(torch ops) vs unpadded(torch ops) {rel(l1t, l0t)}")
    print(f"D={head_dim}: padded(kernels) vs unpadded(torch ops) {rel(l1, l0t)}")
    print(f"D={head_dim}: padded(kernels) vs padded(torch ops) {rel(l1, l1t)}", flush=True)
