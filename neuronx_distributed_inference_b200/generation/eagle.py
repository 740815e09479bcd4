"""EAGLE / EAGLE-3 fused speculation (chain and static token tree).

reference: ``NeuronFusedSpecModel`` EAGLE paths — ``_eagle_context_encoding_forward`` (models/model_base.py:2033-2092),
``_eagle_token_gen_forward`` (:2517-2754), ``_eagle_tree_token_gen_forward`` (:2094-2515), and the rolling hidden-state
buffer (modules/eagle/hidden_state.py:75-161).

An EAGLE draft predicts from *features*: its input at cache slot ``x`` is the pair (token at position ``x+1``, target
feature at position ``x``).  After each verification the last ``k`` accepted pairs are re-fed with the TRUE target features
(read back from the rolling buffers), so draft KV entries computed from the draft's own guessed features never survive a
step.  Everything — the k-wide refresh, the draft chain / tree levels, the target verify, acceptance, KV compaction and the
buffer updates — is tensor code with static shapes and no host synchronisation, so one step is one CUDA-graph replay.

Slots and positions for a step whose root token ``last_token`` sits at position ``p``:
  draft:   root pair at slot p-1; chain token i (tree node n) at slot p-1+i (p-1+n), rotary position p-1+depth
  target:  root at slot p; node n at slot p+n, rotary position p+depth; accepted nodes are compacted to p..p+n_acc-1
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..modules.eagle.hidden_state import HiddenStateRollingBuffer, TokenRollingBuffer
from ..modules.eagle.token_tree import TokenTree
from .speculative import _last_tokens, greedy_accept


class EagleSpeculativeModel(nn.Module):
    def __init__(self, target: nn.Module, draft: nn.Module, speculation_length: int, max_batch_size: int,
                 token_tree: Optional[TokenTree] = None):
        super().__init__()
        self.target_model, self.draft_model = target, draft
        self.tree = token_tree
        self.k = speculation_length if token_tree is None else max(token_tree.max_depth + 1, 2)
        tcfg = target.config
        is3 = bool(target.neuron_config.is_eagle3)
        self.feat_width = tcfg.hidden_size * (3 if is3 else 1)
        dev, dt = target.device_, target.neuron_config.torch_dtype
        L = 2 * max(self.k, token_tree.num_nodes if token_tree is not None else 0)
        self.feat_buf = HiddenStateRollingBuffer(max_batch_size, L, self.feat_width, dt, dev)
        self.tok_buf = TokenRollingBuffer(max_batch_size, L, dev)
        if token_tree is not None:
            token_tree.to(dev)
            # per-level index tensors, created once (no host->device copies inside the step: it is graph captured)
            self._lvl = []
            for d in range(1, token_tree.max_depth + 1):
                nodes = token_tree.level_nodes[d]
                first_prev = token_tree.level_nodes[d - 1][0]
                self._lvl.append(dict(nodes=torch.tensor(nodes, device=dev), par=torch.tensor([token_tree.parent[n] - first_prev for n in nodes], device=dev),
                                      rank=torch.tensor([token_tree.child_rank[n] for n in nodes], device=dev)))
            self._pos_off = token_tree.position_offsets.to(dev)

    def reset(self):
        self.target_model.reset()
        self.draft_model.reset()
        self.feat_buf.reset()
        self.tok_buf.reset()

    def _draft_tokens(self, out):
        return self.draft_model.map_draft_tokens(_last_tokens(out)) if hasattr(self.draft_model, "map_draft_tokens") \
            else _last_tokens(out)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params=None, **kw):
        """Target context encoding exporting every position's feature, then draft context encoding on the shifted pairs
        (token j+1, feature j); the newly sampled token closes the last pair."""
        B, T = input_ids.shape
        dev = input_ids.device
        out_t = self.target_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, is_prefill=True,
                                  output_hidden=True, all_hidden=True, **kw)
        feats = out_t.hidden_states                                              # [B,T,F]
        tok = _last_tokens(out_t).view(B, -1)[:, -1:]
        n = attention_mask.long().sum(-1, keepdim=True)                          # [B,1]
        shifted = torch.cat([input_ids[:, 1:], input_ids[:, -1:]], 1).scatter(1, (n - 1).clamp_min(0), tok.to(input_ids.dtype))
        self.draft_model(shifted, attention_mask, position_ids, seq_ids, None, is_prefill=True, prev_hidden=feats)
        # seed the rings with the last k (feature, token) pairs of the prompt
        back = torch.arange(self.k, device=dev).view(1, -1)
        fpos = n - 1 - back                                                      # feature positions n-1 .. n-k
        ok = fpos >= 0
        g = fpos.clamp_min(0)
        self.feat_buf.set_state_(seq_ids, torch.where(ok, fpos, torch.full_like(fpos, -1)),
                                 feats.gather(1, g.unsqueeze(-1).expand(B, self.k, feats.shape[-1])))
        self.tok_buf.set_tokens(seq_ids, torch.where(ok, fpos + 1, torch.full_like(fpos, -1)), shifted.gather(1, g).long())
        return out_t

    def _refresh_draft(self, last_token, position, seq_ids, want_logits: bool):
        """Re-feed the last k accepted (token, target feature) pairs; returns the draft output at the root pair."""
        k, dev = self.k, last_token.device
        fpos = position - k + torch.arange(k, device=dev, dtype=position.dtype).view(1, k)    # p-k .. p-1
        ok = fpos >= 0
        feats = self.feat_buf.get_state(seq_ids, torch.where(ok, fpos, torch.full_like(fpos, -1)))
        toks = self.tok_buf.get_tokens(seq_ids, torch.where(ok, fpos + 1, torch.full_like(fpos, -1)))
        toks = torch.cat([toks[:, :-1], last_token.long()], 1)
        out = self.draft_model(toks, None, fpos.clamp_min(0), seq_ids, None, is_prefill=False, all_positions=True,
                               prev_hidden=feats, output_hidden=True, output_logits=want_logits,
                               write_positions=torch.where(ok, fpos, torch.full_like(fpos, -1)))
        return out

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, last_token: torch.Tensor, position: torch.Tensor, seq_ids: torch.Tensor, prev_token=None):
        if self.tree is not None:
            return self._tree_step(last_token, position, seq_ids)
        k, B, dev = self.k, last_token.shape[0], last_token.device
        out = self._refresh_draft(last_token, position, seq_ids, False)
        tok = self._draft_tokens(out)[:, -1:]
        feat = out.hidden_states[:, -1:]
        cand = [last_token.long(), tok]
        for i in range(1, k - 1):
            out = self.draft_model(tok, None, position - 1 + i, seq_ids, None, is_prefill=False, prev_hidden=feat, output_hidden=True)
            tok = self._draft_tokens(out).view(B, -1)[:, -1:]
            feat = out.hidden_states[:, -1:]
            cand.append(tok)
        cand_ids = torch.cat(cand[:k], 1)
        cand_pos = position + torch.arange(k, device=dev, dtype=position.dtype).view(1, k)
        out_t = self.target_model(cand_ids, None, cand_pos, seq_ids, None, is_prefill=False, all_positions=True,
                                  output_hidden=True)
        tgt = _last_tokens(out_t).view(B, k)
        accepted, n_acc = greedy_accept(cand_ids[:, 1:], tgt)
        self.feat_buf.set_state_(seq_ids, cand_pos, out_t.hidden_states)
        self.tok_buf.set_tokens(seq_ids, cand_pos + 1, tgt.long())
        next_token = tgt.gather(1, (n_acc - 1).view(B, 1))
        next_pos = position + n_acc.view(B, 1).to(position.dtype)
        return accepted, n_acc, next_token, next_pos, last_token, out_t

    # ------------------------------------------------------------------------------------------------------------
    def _tree_step(self, last_token, position, seq_ids):
        tree, B, dev = self.tree, last_token.shape[0], last_token.device
        N = tree.num_nodes
        root_slot = position - 1                                                  # draft slot of the root pair
        out = self._refresh_draft(last_token, position, seq_ids, True)
        logits = out.logits[:, -1:].float()                                      # [B,1,V] proposals of the root
        feats = out.hidden_states[:, -1:]                                        # [B,1,H]
        cand = torch.zeros(B, N, dtype=torch.long, device=dev)
        cand[:, 0] = last_token.view(B)
        mask_all = tree.attn_mask.to(dev)
        for d in range(1, tree.max_depth + 1):
            nodes = tree.level_nodes[d]
            lv = self._lvl[d - 1]
            par, rank = lv["par"], lv["rank"]
            topk = logits.topk(max(tree.max_children[d - 1], 1), -1).indices      # [B,W_{d-1},K]
            toks = topk[:, par, rank]                                             # [B,W_d]
            if hasattr(self.draft_model, "map_draft_tokens"):
                toks = self.draft_model.map_draft_tokens(toks)
            cand[:, lv["nodes"]] = toks
            if d == tree.max_depth:
                break
            node_t = lv["nodes"].to(position.dtype).view(1, -1)
            out = self.draft_model(toks, None, (root_slot + d).expand(B, len(nodes)), seq_ids, None, is_prefill=False,
                                   all_positions=True, prev_hidden=feats[:, par], output_hidden=True, output_logits=True,
                                   write_positions=root_slot + node_t, active_mask=mask_all[lv["nodes"]].unsqueeze(0),
                                   active_base=root_slot)
            logits, feats = out.logits.float(), out.hidden_states
        node_ids = torch.arange(N, device=dev, dtype=position.dtype).view(1, N)
        out_t = self.target_model(cand, None, position + self._pos_off.view(1, N).to(position.dtype), seq_ids,
                                  None, is_prefill=False, all_positions=True, output_hidden=True,
                                  write_positions=position + node_ids, active_mask=mask_all.unsqueeze(0), active_base=position)
        tgt = _last_tokens(out_t).view(B, N)
        path, n_acc, acc_tok = tree.accept(cand, tgt)                            # [B,L], [B], [B,L]
        L = path.shape[1]
        ar = torch.arange(L, device=dev, dtype=position.dtype).view(1, L)
        keep = path >= 0
        dst = torch.where(keep, position + ar, torch.full_like(path, -1).to(position.dtype))
        src = torch.where(keep, position + path.to(position.dtype), torch.full_like(dst, -1))
        self.target_model.kv_mgr.move(seq_ids, src[:, 1:], dst[:, 1:])           # root stays where it is
        F = out_t.hidden_states
        self.feat_buf.set_state_(seq_ids, dst, F.gather(1, path.clamp_min(0).unsqueeze(-1).expand(B, L, F.shape[-1])))
        self.tok_buf.set_tokens(seq_ids, torch.where(keep, dst + 1, dst), acc_tok.clamp_min(0))
        next_token = acc_tok.gather(1, (n_acc - 1).view(B, 1))
        next_pos = position + n_acc.view(B, 1).to(position.dtype)
        return acc_tok, n_acc, next_token, next_pos, last_token, out_t
