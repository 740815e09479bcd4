"""Dynamic token trees (EAGLE-2 style): the tree shape is chosen per step from the draft's own confidence.

reference: modules/eagle/dynamic_token_tree.py:4-352 (adjacency updates per draft step, cumulative draft probability per
node, selection of the ``num_verify`` best nodes and their paths).

Each draft step expands the ``step_width`` most probable frontier nodes by ``branching_factor`` children; every node
carries the cumulative log-probability of its path.  After ``steps`` expansions the ``num_verify - 1`` best non-root nodes
(closed under ancestors, because a child's cumulative probability never exceeds its parent's) form the verification tree.
All bookkeeping is batched tensor code with static shapes: [B, max_nodes]."""
from __future__ import annotations

from typing import Tuple

import torch


class DynamicTokenTree:
    def __init__(self, steps: int, branching_factor: int, step_width: int, num_verify: int):
        self.steps, self.branch, self.width, self.num_verify = steps, branching_factor, step_width, num_verify
        # node 0 root; step 0 adds `branch` nodes (children of root); later steps add width*branch nodes
        self.nodes_per_step = [branching_factor] + [step_width * branching_factor] * (steps - 1)
        self.max_nodes = 1 + sum(self.nodes_per_step)

    def init_state(self, root_token: torch.Tensor):
        B, dev = root_token.shape[0], root_token.device
        N = self.max_nodes
        st = dict(tokens=torch.zeros(B, N, dtype=torch.long, device=dev),
                  parent=torch.full((B, N), -1, dtype=torch.long, device=dev),
                  depth=torch.zeros(B, N, dtype=torch.long, device=dev),
                  score=torch.full((B, N), float("-inf"), device=dev),
                  n=1)
        st["tokens"][:, 0] = root_token
        st["score"][:, 0] = 0.0
        return st

    def expand(self, st, frontier: torch.Tensor, logprobs: torch.Tensor) -> torch.Tensor:
        """frontier [B,W] node ids that were just run through the draft; logprobs [B,W,V] their next-token log-probs.
        Adds W*branch children; returns the next frontier [B, step_width] (the best new nodes)."""
        B, W, _ = logprobs.shape
        top_lp, top_tok = logprobs.topk(self.branch, -1)                        # [B,W,b]
        base = st["score"].gather(1, frontier).unsqueeze(-1)                    # [B,W,1]
        new_score = (base + top_lp).reshape(B, -1)
        new_tok = top_tok.reshape(B, -1)
        new_parent = frontier.unsqueeze(-1).expand(B, W, self.branch).reshape(B, -1)
        new_depth = st["depth"].gather(1, new_parent) + 1
        n0, cnt = st["n"], W * self.branch
        sl = slice(n0, n0 + cnt)
        st["tokens"][:, sl], st["parent"][:, sl], st["depth"][:, sl], st["score"][:, sl] = new_tok, new_parent, new_depth, new_score
        st["n"] = n0 + cnt
        k = min(self.width, cnt)
        return new_score.topk(k, -1).indices + n0

    def ancestor_mask(self, st) -> torch.Tensor:
        """[B,N,N] node i sees node j iff j is i or an ancestor of i."""
        B, N = st["parent"].shape
        dev = st["parent"].device
        m = torch.eye(N, dtype=torch.bool, device=dev).unsqueeze(0).repeat(B, 1, 1)
        cur = st["parent"].clone()
        for _ in range(self.steps + 1):
            ok = cur >= 0
            m |= torch.nn.functional.one_hot(cur.clamp_min(0), N).bool() & ok.unsqueeze(-1)
            cur = torch.where(ok, st["parent"].gather(1, cur.clamp_min(0)), cur)
        return m

    def select(self, st) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Pick the verification tree: -> (node_ids [B,M] sorted by (depth, id) with the root first, tokens [B,M],
        depth [B,M], mask [B,M,M])."""
        M = min(self.num_verify, st["n"])
        score = st["score"].clone()
        score[:, 0] = float("inf")
        sel = score.topk(M, -1).indices
        key = st["depth"].gather(1, sel) * self.max_nodes + sel
        sel = sel.gather(1, key.argsort(-1))
        full = self.ancestor_mask(st)
        B = sel.shape[0]
        mask = full[torch.arange(B, device=sel.device).view(B, 1, 1), sel.unsqueeze(-1), sel.unsqueeze(1)]
        return sel, st["tokens"].gather(1, sel), st["depth"].gather(1, sel), mask

    @staticmethod
    def accept(tokens: torch.Tensor, depth: torch.Tensor, mask: torch.Tensor, target_tokens: torch.Tensor):
        """Greedy acceptance on a per-row tree given as an ancestor mask.  tokens/depth/target_tokens [B,M], mask [B,M,M].
        A node is *consistent* iff its token equals the target's prediction after its parent; a node is accepted iff all
        its ancestors and itself are consistent.  -> (path [B,Dmax+1] local indices padded -1, n_acc [B], acc_tok)."""
        B, M = tokens.shape
        dev = tokens.device
        anc = mask & ~torch.eye(M, dtype=torch.bool, device=dev)
        # parent = the ancestor of greatest depth
        pd = torch.where(anc, depth.unsqueeze(1).expand(B, M, M), torch.full((B, M, M), -1, device=dev, dtype=depth.dtype))
        parent = pd.argmax(-1)
        has_parent = anc.any(-1)
        consistent = (tokens == target_tokens.gather(1, parent)) | ~has_parent
        all_ok = ((~mask) | consistent.unsqueeze(1)).all(-1)                    # every visible node consistent
        score = torch.where(all_ok, depth, torch.full_like(depth, -1))
        leaf = score.argmax(-1)                                                  # deepest accepted node
        n_acc = score.max(-1).values + 1
        D = int(depth.max()) + 1
        on_path = mask[torch.arange(B, device=dev), leaf]                        # [B,M]
        order = torch.where(on_path, depth, torch.full_like(depth, M + D))
        path = order.argsort(-1)[:, :D]
        ar = torch.arange(D, device=dev).view(1, D)
        keep = ar < n_acc.view(B, 1)
        path = torch.where(keep, path, torch.full_like(path, -1))
        acc_tok = torch.where(keep, target_tokens.gather(1, path.clamp_min(0)), torch.full_like(path, -1))
        return path, n_acc, acc_tok
