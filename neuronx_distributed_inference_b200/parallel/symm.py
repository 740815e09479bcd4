"""Symmetric (peer-mapped) workspace for the in-kernel collectives.

Every rank of a group ``cudaMalloc``s the same buffers, exports them through CUDA IPC, and
all-gathers the handles over ``torch.distributed`` (NCCL is only the bootstrap here); afterwards each
rank holds a table of peer-mapped device pointers and the fused kernels store straight into
peer HBM over NVLink (csrc/gemv2.cu MODE 1).

Protocol of the one-shot fused all-reduce (T <= 8 tokens, the decode RowParallel path), "LL" style:
    recv  [2 parity][world src][8 tok][n_max] x {fp32 value, u32 tag}
  * sender: every finished output element travels as ONE 8-byte ``{value, tag}`` store into slot
    [parity][me] of EVERY rank (no separate flag, no fence.sys round trip);
  * receiver: polls the ``world`` slots of an element together until all carry the expected tag, sums the
    sources in rank order (bitwise identical result on every rank), adds bias + residual, writes bf16;
  * tag = (step << 8 | call) + 1: ``step`` is a DEVICE counter bumped by ``begin_step()`` at the start of
    every forward (the bump is an in-stream op, so CUDA-graph replays see fresh tags), ``call`` the index of
    the collective inside the forward.  A skipped / repeated collective therefore times out and traps
    instead of silently consuming stale data, and slots need no reset;
  * parity alternates per call on the host (double buffer: a rank can be at most one collective ahead of
    its slowest peer); graphs contain an even number of calls (the runner pads with one dummy call
    otherwise) so that replays and eager calls stay in phase.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops._ext import load_extension

MAX_TOKENS = 8


class SymmetricWorkspace:
    ARGMAX_ROWS = 1024

    def __init__(self, group, device: torch.device, n_max: int):
        self.group = group
        self.device = device
        self.world = group.size
        self.rank = group.rank
        self.n_max = int(n_max)
        self.parity = 0
        self.calls = 0          # collectives issued so far (host-side, diagnostics)
        self.call = 0           # index of the next collective inside the current forward (part of the tag)
        self.step_t = torch.zeros(1, dtype=torch.int32, device=device)   # device-side step counter (part of the tag)
        self._C = load_extension()
        recv_bytes = 2 * self.world * MAX_TOKENS * self.n_max * 8   # {fp32 value, u32 tag} per element (LL)
        self._local_recv, h_recv = self._C.symm_alloc(recv_bytes)
        # arg-max exchange slots: [2 parity][ARGMAX_ROWS][world src] x {value, tag, global index, tag}
        self._local_slots, h_slots = self._C.symm_alloc(2 * self.ARGMAX_ROWS * self.world * 16)
        handles: List = [None] * self.world
        dist.all_gather_object(handles, (bytes(h_recv), bytes(h_slots)), group=group.pg)
        self.recv_ptrs, self.slot_ptrs, self._opened = [], [], []
        for r, (hr, hs) in enumerate(handles):
            if r == self.rank:
                self.recv_ptrs.append(self._local_recv)
                self.slot_ptrs.append(self._local_slots)
            else:
                pr, ps = self._C.symm_open(hr), self._C.symm_open(hs)
                self._opened += [pr, ps]
                self.recv_ptrs.append(pr)
                self.slot_ptrs.append(ps)
        torch.cuda.synchronize(device)
        dist.barrier(group=group.pg)

    @classmethod
    def create(cls, group, device, max_tokens: int = MAX_TOKENS, max_width: int = 8192) -> "SymmetricWorkspace":
        assert max_tokens <= MAX_TOKENS
        n_max = max(int(max_width), 1024)
        n_max = (n_max + 15) // 16 * 16
        return cls(group, device, n_max)

    def begin_step(self):
        """Start of a forward (eager or being captured): bump the device step counter, restart the call index."""
        self.step_t.add_(1)
        self.call = 0

    def gemv_allreduce(self, x, w, bias=None, residual=None, scale=None):
        """``scale``: fp32 per-output-channel (or per-tensor) dequantisation scale of int8 / fp8 weights."""
        y = self._C.gemv_allreduce(x, w, bias, residual, self.recv_ptrs, self.step_t, self.rank, self.parity, self.call,
                                   self.n_max, scale)
        self.parity ^= 1
        self.call += 1
        self.calls += 1
        return y

    def argmax(self, logits):
        """Vocabulary-sharded arg-max with the cross-rank exchange inside the kernel; counts as one collective."""
        y = self._C.argmax(logits, self.slot_ptrs, self.step_t, self.rank, self.parity, self.call, self.ARGMAX_ROWS)
        self.parity ^= 1
        self.call += 1
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
