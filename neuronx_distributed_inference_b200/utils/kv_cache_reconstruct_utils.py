"""KV-cache reconstruction: un-shard the per-rank device caches into the Hugging Face layout ``[B, H_kv, S, D]`` so they
can be compared with golden ``past_key_values`` (reference utils/kv_cache_reconstruct_utils.py:57-251, KVCacheReconstructREADME.md).
Undoes, per layer: TP head sharding (including REPLICATE_TO_TP_DEGREE duplicates and CONVERT_TO_MHA expansion through the
GQA index plan), the garbage line, fp8 storage, and cache-line (seq_id) permutation."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from ..parallel import mappings
from ..parallel.state import get_tensor_model_parallel_group


def reconstruct_kv_cache(app, seq_ids: Optional[torch.Tensor] = None, seq_len: Optional[int] = None,
                         gather: bool = True) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    model = app.model
    mgr = model.kv_mgr
    g = get_tensor_model_parallel_group()
    out = []
    for i, layer in enumerate(model.layers):
        k, v = mgr.get_kv_by_layer_id(i)
        k, v = k[: mgr.num_lines], v[: mgr.num_lines]
        if k.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            k = k.float() * (mgr.k_scale or 1.0)
            v = v.float() * (mgr.v_scale or 1.0)
        if seq_ids is not None:
            k, v = k[seq_ids.long()], v[seq_ids.long()]
        if seq_len is not None:
            k, v = k[:, :, :seq_len], v[:, :, :seq_len]
        attn = getattr(layer, "self_attn", None)
        plan = getattr(getattr(attn, "qkv_proj", None), "plan", None)
        if g.size > 1 and gather:
            ks = mappings.all_gather(k.unsqueeze(0).contiguous(), 0, g)     # [tp, B, H_loc, S, D]
            vs = mappings.all_gather(v.unsqueeze(0).contiguous(), 0, g)
            if plan is not None:
                n_kv = plan.n_kv
                kk = k.new_zeros((k.shape[0], n_kv) + tuple(k.shape[2:]))
                vv = v.new_zeros((v.shape[0], n_kv) + tuple(v.shape[2:]))
                for r in range(g.size):
                    for j, src in enumerate(plan.kv_idx[r]):
                        kk[:, src] = ks[r][:, j]
                        vv[:, src] = vs[r][:, j]
                k, v = kk, vv
            else:
                k = ks.permute(1, 0, 2, 3, 4).reshape(k.shape[0], -1, *k.shape[2:])
                v = vs.permute(1, 0, 2, 3, 4).reshape(v.shape[0], -1, *v.shape[2:])
        elif plan is not None and plan.kv_per_rank != plan.n_kv:
            # single rank but heads were duplicated (convert-to-MHA): keep the first copy of each source head
            n_kv = plan.n_kv
            first = {}
            for j, src in enumerate(plan.kv_idx[g.rank]):
                first.setdefault(src, j)
            sel = torch.tensor([first[s] for s in range(n_kv)], device=k.device)
            k, v = k[:, sel], v[:, sel]
        out.append((k.float().cpu(), v.float().cpu()))
    return out


def compare_kv_cache(reconstructed, golden, rtol=1e-2, atol=1e-3, seq_len: Optional[int] = None):
    """golden: HF ``past_key_values`` (iterable of (k, v) ``[B,H,S,D]``).  -> list of (layer, max_abs_err_k, max_abs_err_v, ok)."""
    rep = []
    for i, ((k, v), gv) in enumerate(zip(reconstructed, golden)):
        gk, gvv = gv[0].float().cpu(), gv[1].float().cpu()
        S = min(k.shape[2], gk.shape[2]) if seq_len is None else seq_len
        ek = (k[:, :, :S] - gk[:, :, :S]).abs()
        evv = (v[:, :, :S] - gvv[:, :, :S]).abs()
        ok = bool((ek <= atol + rtol * gk[:, :, :S].abs()).all() and (evv <= atol + rtol * gvv[:, :, :S].abs()).all())
        rep.append((i, float(ek.max()), float(evv.max()), ok))
    return rep
