"""Static token trees for tree speculation (EAGLE token-tree / Medusa).

reference: modules/eagle/token_tree.py:8-646 — a tree is given as a JSON adjacency dict ``{"0": ["1","2"], "1": ["3"], ...}``
(node 0 = root = the last accepted token); from it the reference derives per-level attention masks, root-to-leaf paths,
rotary position offsets, cache scatter indices and permutation indices that are threaded through the traced graph.

Here the tree is reduced to the index tensors the fused step needs:

* ``parent[n]``, ``depth[n]``, ``child_rank[n]`` (node n is the ``child_rank``-th most likely continuation of its parent);
* ``level_nodes[d]`` — nodes of depth d (BFS order == node order, so a level is a contiguous range);
* ``attn_mask [N,N]`` — node i sees node j iff j is i or an ancestor of i (visibility among active tokens);
* ``position_offsets [N]`` — rotary position of node n relative to the root (= depth), while the KV *slot* offset is n;
* ``paths [P, max_depth+1]`` — every root-to-leaf path, padded with -1 (acceptance = longest matching path prefix);
* ``level_width``, ``max_children`` per level (the top-k width the draft needs at that level).

Nodes are renumbered in BFS order, so any consistent labelling of the JSON is accepted."""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Union

import torch


def _load(tree: Union[str, Dict, Sequence]) -> Dict[int, List[int]]:
    if isinstance(tree, str):
        with open(tree) as f:
            tree = json.load(f)
    if isinstance(tree, dict):
        return {int(k): [int(c) for c in v] for k, v in tree.items()}
    # Medusa style: list of paths in child-rank coordinates, e.g. [[0],[0,0],[1],[0,1]]
    adj: Dict[int, List[int]] = {0: []}
    ids = {(): 0}
    for path in sorted([tuple(p) for p in tree], key=lambda p: (len(p), p)):
        for d in range(1, len(path) + 1):
            pre = path[:d]
            if pre not in ids:
                ids[pre] = len(ids)
                adj.setdefault(ids[pre[:-1]], []).append(ids[pre])
                adj.setdefault(ids[pre], [])
    # children must be ordered by their rank coordinate
    inv = {v: k for k, v in ids.items()}
    for k in adj:
        adj[k].sort(key=lambda c: inv[c][-1])
    adj["__rank__"] = {ids[p]: p[-1] for p in ids if p}   # type: ignore
    return adj


class TokenTree:
    def __init__(self, tree_config: Union[str, Dict, Sequence]):
        adj = _load(tree_config)
        explicit_rank = adj.pop("__rank__", None)
        children_all = {c for v in adj.values() for c in v}
        roots = [k for k in adj if k not in children_all]
        if len(roots) != 1:
            raise ValueError(f"token tree needs exactly one root, found {roots}")
        # BFS renumbering
        order, parent_old = [roots[0]], {roots[0]: -1}
        i = 0
        while i < len(order):
            for c in adj.get(order[i], []):
                if c in parent_old:
                    raise ValueError(f"node {c} has two parents")
                parent_old[c] = order[i]
                order.append(c)
            i += 1
        new = {o: n for n, o in enumerate(order)}
        N = len(order)
        self.num_nodes = N
        self.parent = [-1] * N
        self.child_rank = [0] * N
        self.children: List[List[int]] = [[] for _ in range(N)]
        for o in order:
            for r, c in enumerate(adj.get(o, [])):
                self.parent[new[c]] = new[o]
                self.child_rank[new[c]] = explicit_rank[c] if explicit_rank else r
                self.children[new[o]].append(new[c])
        self.depth = [0] * N
        for n in range(1, N):
            self.depth[n] = self.depth[self.parent[n]] + 1
        self.max_depth = max(self.depth)
        self.level_nodes = [[n for n in range(N) if self.depth[n] == d] for d in range(self.max_depth + 1)]
        self.level_width = [len(l) for l in self.level_nodes]
        self.max_children = [max((max((self.child_rank[c] for c in self.children[n]), default=-1) + 1
                                  for n in self.level_nodes[d]), default=0) for d in range(self.max_depth + 1)]
        m = torch.zeros(N, N, dtype=torch.bool)
        for n in range(N):
            a = n
            while a >= 0:
                m[n, a] = True
                a = self.parent[a]
        self.attn_mask = m
        self.position_offsets = torch.tensor(self.depth, dtype=torch.int32)
        leaves = [n for n in range(N) if not self.children[n]]
        paths = []
        for leaf in leaves:
            p, a = [], leaf
            while a >= 0:
                p.append(a)
                a = self.parent[a]
            p = p[::-1]
            paths.append(p + [-1] * (self.max_depth + 1 - len(p)))
        self.paths = torch.tensor(paths, dtype=torch.long)
        self.parent_t = torch.tensor(self.parent, dtype=torch.long)
        self.child_rank_t = torch.tensor(self.child_rank, dtype=torch.long)

    # ---- derived masks (reference token_tree.py level masks) ----
    def level_mask(self, d: int) -> torch.Tensor:
        """[width_d, N] visibility of the nodes of level d."""
        return self.attn_mask[self.level_nodes[d]]

    def to(self, device):
        for k in ("attn_mask", "position_offsets", "paths", "parent_t", "child_rank_t"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def fill_candidates(self, root_token: torch.Tensor, level_topk: List[torch.Tensor]) -> torch.Tensor:
        """root_token [B]; level_topk[d] = [B, width_d, K] top-K continuations proposed *by* the nodes of level d.
        -> candidate token per node [B,N]."""
        B = root_token.shape[0]
        cand = torch.zeros(B, self.num_nodes, dtype=torch.long, device=root_token.device)
        cand[:, 0] = root_token
        for d in range(1, self.max_depth + 1):
            prev = self.level_nodes[d - 1]
            first_prev = prev[0]
            for n in self.level_nodes[d]:
                cand[:, n] = level_topk[d - 1][:, self.parent[n] - first_prev, self.child_rank[n]]
        return cand

    def accept(self, cand: torch.Tensor, target_tokens: torch.Tensor):
        """Greedy tree acceptance.  cand [B,N] (node tokens), target_tokens [B,N] (target arg-max *after* each node).
        -> (best_path [B, max_depth+1] node ids padded -1, n_acc [B] accepted nodes incl. root in 1..max_depth+1,
            accepted tokens [B, max_depth+1] = target tokens along the path, padded -1)."""
        B = cand.shape[0]
        paths = self.paths.to(cand.device)                       # [P, L]
        P, L = paths.shape
        valid = paths >= 0
        pc = paths.clamp_min(0)
        node_tok = cand[:, pc]                                   # [B,P,L] token at each path node
        tgt_prev = target_tokens[:, pc]                          # [B,P,L] target's prediction after each node
        # node at depth d (d>=1) is accepted iff its token equals the target's prediction after its parent (depth d-1)
        ok = torch.ones(B, P, L, dtype=torch.bool, device=cand.device)
        ok[:, :, 1:] = (node_tok[:, :, 1:] == tgt_prev[:, :, :-1]) & valid[None, :, 1:]
        n_match = ok.long().cumprod(-1).sum(-1)                  # [B,P] accepted nodes incl. root
        n_acc, best = n_match.max(-1)
        best_path = paths[best]                                  # [B,L]
        ar = torch.arange(L, device=cand.device).view(1, L)
        keep = ar < n_acc.view(B, 1)
        best_path = torch.where(keep, best_path, torch.full_like(best_path, -1))
        acc_tok = torch.where(keep, target_tokens.gather(1, best_path.clamp_min(0)), torch.full_like(best_path, -1))
        return best_path, n_acc, acc_tok
