"""Medusa speculation: extra decoding heads on the target propose a token tree, one target forward verifies it.

reference: ``NeuronBaseModel._medusa_forward`` (models/model_base.py:393-508), the host loop
``_medusa_assisted_decoding`` (utils/hf_adapter.py:799-915) and the tree buffers it builds from ``medusa_tree``
(``generate_medusa_buffers`` / ``generate_candidates`` / ``evaluate_posterior`` / ``update_inference_inputs``).

The reference verifies on the device and does candidate generation / acceptance / cache bookkeeping on the host with
``accepted_indices`` + ``current_length`` scatter tensors.  Here the whole step is device code: candidates from the head
top-k, one tree-masked target forward, greedy path acceptance (:class:`TokenTree`), in-place KV compaction."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from ..modules.eagle.token_tree import TokenTree

DEFAULT_MEDUSA_TREE = [[0], [0, 0], [1], [0, 1], [2], [0, 0, 0], [1, 0], [0, 2], [3], [0, 3], [4], [0, 4], [2, 0], [0, 5],
                       [0, 0, 1], [5], [0, 6], [6], [0, 7], [0, 1, 0], [1, 1], [7], [0, 8], [0, 0, 2], [3, 0], [0, 9], [8],
                       [9], [1, 0, 0], [0, 2, 0], [1, 2], [0, 0, 3], [4, 0], [2, 1], [0, 0, 4], [0, 0, 5], [0, 0, 0, 0],
                       [0, 1, 1], [0, 0, 6], [0, 3, 0], [5, 0], [1, 3], [0, 0, 7], [0, 0, 8], [0, 0, 9], [6, 0], [0, 4, 0],
                       [1, 4], [7, 0], [0, 1, 2], [2, 0, 0], [3, 1], [2, 2], [8, 0], [0, 5, 0], [1, 5], [1, 0, 1], [0, 2, 1],
                       [9, 0], [0, 6, 0], [0, 0, 0, 1], [1, 6], [0, 7, 0]]


def generate_medusa_buffers(medusa_tree: List[List[int]], device="cpu"):
    """Same dictionary the reference's host loop consumes (attention mask, tree indices into the flattened head top-k,
    position ids, retrieve indices = root-to-leaf paths)."""
    t = TokenTree(medusa_tree)
    topk = max(t.max_children) if t.max_children else 1
    tree_indices = torch.zeros(t.num_nodes, dtype=torch.long)
    for n in range(1, t.num_nodes):
        tree_indices[n] = 1 + (t.depth[n] - 1) * topk + t.child_rank[n]
    return {"medusa_attn_mask": t.attn_mask.float().to(device), "tree_indices": tree_indices.to(device),
            "medusa_position_ids": t.position_offsets.long().to(device), "retrieve_indices": t.paths.to(device), "tree": t,
            "topk": topk}


class MedusaSpeculativeModel(nn.Module):
    def __init__(self, target: nn.Module, medusa_tree=None):
        super().__init__()
        self.target_model = target
        self.tree = TokenTree(medusa_tree or DEFAULT_MEDUSA_TREE).to(target.device_)
        nh = target.neuron_config.num_medusa_heads
        if self.tree.max_depth > nh:
            raise ValueError(f"medusa tree depth {self.tree.max_depth} exceeds num_medusa_heads {nh}")
        self.K = max(max(self.tree.max_children), 1)
        dev = target.device_
        self._depth = torch.tensor(self.tree.depth, device=dev)
        self._rank = self.tree.child_rank_t.to(dev)

    def reset(self):
        self.target_model.reset()

    def _heads_topk(self, medusa_logits, index):
        """medusa_logits [Hd,B,T,V]; index [B] position along T -> [B,Hd,K]."""
        Hd, B = medusa_logits.shape[:2]
        sel = medusa_logits[:, torch.arange(B, device=index.device), index]      # [Hd,B,V]
        return sel.float().topk(self.K, -1).indices.permute(1, 0, 2)

    @torch.no_grad()
    def prefill(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params=None, **kw):
        out = self.target_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, is_prefill=True,
                                output_logits=True, **kw)
        B = input_ids.shape[0]
        root = out.logits[:, -1].argmax(-1)
        heads = self._heads_topk(out.extras["medusa_logits"], torch.zeros(B, dtype=torch.long, device=root.device))
        return out, root, heads

    @torch.no_grad()
    def forward(self, root_token: torch.Tensor, head_topk: torch.Tensor, position: torch.Tensor, seq_ids: torch.Tensor):
        """root_token [B] sits at ``position`` [B,1]; head_topk [B,Hd,K] are the head proposals made together with it.
        -> (accepted tokens [B,L] padded -1, n_acc [B], next_root [B], next_head_topk, next_position)."""
        tree, B, dev = self.tree, root_token.shape[0], root_token.device
        N = tree.num_nodes
        depth = self._depth
        cand = torch.empty(B, N, dtype=torch.long, device=dev)
        cand[:, 0] = root_token
        if N > 1:
            cand[:, 1:] = head_topk[:, depth[1:] - 1, self._rank[1:]]
        node_ids = torch.arange(N, device=dev, dtype=position.dtype).view(1, N)
        out = self.target_model(cand, None, position + depth.view(1, N).to(position.dtype), seq_ids, None, is_prefill=False,
                                all_positions=True, output_logits=True, write_positions=position + node_ids,
                                active_mask=tree.attn_mask.unsqueeze(0), active_base=position)
        tgt = out.logits.argmax(-1)
        path, n_acc, acc_tok = tree.accept(cand, tgt)
        L = path.shape[1]
        ar = torch.arange(L, device=dev, dtype=position.dtype).view(1, L)
        keep = path >= 0
        dst = torch.where(keep, position + ar, torch.full_like(ar.expand(B, L), -1))
        src = torch.where(keep, position + path.to(position.dtype), torch.full_like(dst, -1))
        self.target_model.kv_mgr.move(seq_ids, src[:, 1:], dst[:, 1:])
        last_node = path.gather(1, (n_acc - 1).view(B, 1)).view(B)
        next_root = acc_tok.gather(1, (n_acc - 1).view(B, 1)).view(B)
        next_heads = self._heads_topk(out.extras["medusa_logits"], last_node)
        return acc_tok, n_acc, next_root, next_heads, position + n_acc.view(B, 1).to(position.dtype)


@torch.no_grad()
def medusa_generate(adapter, input_ids, attention_mask, max_length, eos: List[int], pad_id: int,
                    return_dict_in_generate: bool = False):
    model = adapter.neuron_model
    med = model.medusa_model
    dev = model.device
    B = input_ids.shape[0]
    model.reset()
    seq_ids = torch.arange(B, dtype=torch.int32, device=dev)
    pos0 = (attention_mask.long().cumsum(-1) - 1).clamp_min(0).to(torch.int32)
    _, root, heads = med.prefill(input_ids.to(dev), attention_mask.to(dev), pos0.to(dev), seq_ids)
    n_valid = attention_mask.sum(-1).view(B, 1).to(dev)
    position = n_valid.to(torch.int32)
    rows = [input_ids[b, : int(n_valid[b])].tolist() + [int(root[b])] for b in range(B)]
    done = [int(root[b]) in eos for b in range(B)]
    stats = {"steps": 0, "accepted": 0}
    from .speculative import maybe_graph_step
    step = maybe_graph_step(med, lambda r, h, s, p: med(r, h, p, s), [root, heads, seq_ids, position], 5, ("medusa", B))
    while not all(done) and min(len(r) for r, d in zip(rows, done) if not d) < max_length:
        acc, n_acc, root, heads, position = step(root, heads, seq_ids, position)
        stats["steps"] += 1
        stats["accepted"] += int(n_acc.sum())
        for b, toks in enumerate(acc.cpu().tolist()):
            for t in toks:
                if done[b] or t < 0:
                    break
                rows[b].append(t)
                if t in eos or len(rows[b]) >= max_length:
                    done[b] = True
    width = max(len(r) for r in rows)
    seqs = torch.full((B, width), pad_id, dtype=torch.long)
    for b, r in enumerate(rows):
        seqs[b, : len(r)] = torch.tensor(r)
    if return_dict_in_generate:
        from ..utils.hf_adapter import GenerateOutput
        return GenerateOutput(sequences=seqs, speculation_stats=stats)
    return seqs
