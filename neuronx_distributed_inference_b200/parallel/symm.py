"""Symmetric (peer-mapped) workspace for the in-kernel collectives.

Every rank of a group ``cudaMalloc``s the same buffers, exports them through CUDA IPC, and
all-gathers the handles over ``torch.distributed`` (NCCL is only the bootstrap here); afterwards each
rank holds a table of peer-mapped device pointers and the fused kernels store / signal straight into
peer HBM over NVLink (csrc/gemv.cuh MODE 1).

Protocol of the one-shot fused all-reduce (T <= 8 tokens, the decode RowParallel path):
    recv  [2 parity][world src][8 tok][n_max] fp32     flags [2 parity][world src][SYMM_MAX_TILES] u32
  * sender: after reducing a 16-column tile across its warps, stores its partial into slot [parity][me]
    of EVERY rank, then (fence.sys, st.release.sys) sets flag [parity][me][tile] = 1 on every rank;
  * receiver: spins (ld.acquire.sys, bounded, traps on timeout) on its own flags of all sources for its
    tiles, resets them to 0, sums the sources in rank order (bitwise identical result on every rank),
    adds bias + residual and writes bf16.
  * parity alternates per call on the host; graphs must contain an even number of calls (the runner
    pads with one dummy call otherwise) so that replays and eager calls stay in phase.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops._ext import load_extension

SYMM_MAX_TILES = 1024
MAX_TOKENS = 8


class SymmetricWorkspace:
    def __init__(self, group, device: torch.device, n_max: int):
        self.group = group
        self.device = device
        self.world = group.size
        self.rank = group.rank
        self.n_max = int(n_max)
        self.parity = 0
        self.calls = 0
        self._C = load_extension()
        recv_bytes = 2 * self.world * MAX_TOKENS * self.n_max * 8   # {fp32 value, u32 flag} per element (LL)
        flag_bytes = 2 * self.world * SYMM_MAX_TILES * 4
        self._local_recv, h_recv = self._C.symm_alloc(recv_bytes)
        self._local_flags, h_flags = self._C.symm_alloc(flag_bytes)
        handles: List = [None] * self.world
        dist.all_gather_object(handles, (bytes(h_recv), bytes(h_flags)), group=group.pg)
        self.recv_ptrs, self.flag_ptrs, self._opened = [], [], []
        for r, (hr, hf) in enumerate(handles):
            if r == self.rank:
                self.recv_ptrs.append(self._local_recv)
                self.flag_ptrs.append(self._local_flags)
            else:
                pr, pf = self._C.symm_open(hr), self._C.symm_open(hf)
                self._opened += [pr, pf]
                self.recv_ptrs.append(pr)
                self.flag_ptrs.append(pf)
        torch.cuda.synchronize(device)
        dist.barrier(group=group.pg)

    @classmethod
    def create(cls, group, device, max_tokens: int = MAX_TOKENS, max_width: int = 8192) -> "SymmetricWorkspace":
        assert max_tokens <= MAX_TOKENS
        n_max = max(int(max_width), 1024)
        n_max = min((n_max + 15) // 16 * 16, SYMM_MAX_TILES * 16)
        return cls(group, device, n_max)

    def gemv_allreduce(self, x, w, bias=None, residual=None, scale=None, nxt=None):
        assert scale is None, "quantised fused all-reduce not wired yet"
        y = self._C.gemv_allreduce(x, w, bias, residual, self.recv_ptrs, self.flag_ptrs, self.rank, self.parity,
                                   self.n_max, nxt[0] if nxt else None, bool(nxt[1]) if nxt else False)
        self.parity ^= 1
        self.calls += 1
        return y

    def _dummy(self):
        if not hasattr(self, "_dz"):
            self._dz = torch.zeros(1, 256, dtype=torch.bfloat16, device=self.device)
            self._dw = torch.zeros(16, 256, dtype=torch.bfloat16, device=self.device)
        self.gemv_allreduce(self._dz, self._dw)

    def ensure_even(self):
        """Bring the host parity back to 0.  Called before a CUDA-graph capture/replay (graphs bake the
        parity sequence 0,1,0,1,... of their collectives) and at the end of a capture, so that every unit of
        work — an eager forward or a graph — starts on parity 0 and strict alternation holds globally."""
        if self.parity == 1:
            self._dummy()

    def close(self):
        for p in self._opened:
            self._C.symm_close(p)
        self._opened = []
