"""Symmetric heap of a tensor-parallel group: one VMM allocation per rank, peer-mapped everywhere, plus the NVLS multicast
mapping of all of them (csrc/symm_heap.cpp), and the in-switch collectives that run on it (csrc/nvls.cu).

Bootstrap: the POSIX file descriptors of the allocations travel between the ranks over Unix-domain sockets (SCM_RIGHTS); the
socket paths are exchanged through ``torch.distributed`` — NCCL / gloo is only the rendezvous.  After construction a rank holds

    local_va            its own allocation
    peer_va[r]          rank r's allocation (NVLink peer mapping)
    mc_va               the multicast address range: stores reach every copy, ``multimem.ld_reduce`` sums all copies in the switch

``alloc`` is a bump allocator: every rank performs the same sequence of calls, so offsets are symmetric.
Layout: [0, 64 KB) signal words of the collectives, then buffers.

Collectives (bf16, prefill-sized; the decode-sized all-reduce is fused into the GEMV, parallel/symm.py):
    all_reduce      two-shot in-switch: rank r pulls the reduced slice r and multicasts it
    reduce_scatter  rank r pulls its reduced slice (sequence-parallel row-parallel layers), residual add fused
    all_gather      rank r multicasts its slice
reference: the collectives of the external ``neuronx_distributed.parallel_layers.mappings``.
"""
from __future__ import annotations

import os
import socket
import tempfile
import uuid
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..ops._ext import load_extension

_SIG_BYTES = 64 * 1024


def _exchange_fds(group, send: Dict[int, int], n_expect: int, tag: str) -> Dict[int, int]:
    """Send ``send[peer] = fd`` to each peer, receive ``n_expect`` descriptors -> {source rank: fd}."""
    rank, world = group.rank, group.size
    path = os.path.join(tempfile.gettempdir(), f"nxdi_heap_{tag}_{rank}.sock")
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(world + 1)
    paths: List = [None] * world
    dist.all_gather_object(paths, path, group=group.pg)      # everyone is listening once this returns
    for peer, fd in send.items():
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(paths[peer])
        socket.send_fds(c, [rank.to_bytes(4, "little")], [fd])
        c.close()
    got: Dict[int, int] = {}
    for _ in range(n_expect):
        conn, _ = srv.accept()
        msg, fds, _, _ = socket.recv_fds(conn, 4, 1)
        got[int.from_bytes(msg, "little")] = fds[0]
        conn.close()
    dist.barrier(group=group.pg)
    srv.close()
    os.unlink(path)
    return got


class SymmetricHeap:
    def __init__(self, group, device: torch.device, nbytes: int = 256 << 20):
        self.C = C = load_extension()
        self.group, self.device = group, device
        self.rank, self.world = group.rank, group.size
        dev = device.index if device.index is not None else torch.cuda.current_device()
        tags: List = [None] * self.world
        dist.all_gather_object(tags, uuid.uuid4().hex[:12], group=group.pg)
        tag = tags[0]
        self.h = C.symm_heap_create(int(nbytes), dev, self.world, self.rank)
        self.size = int(C.symm_heap_size(self.h))
        self.local_va = int(C.symm_heap_local_va(self.h))
        # ---- peer mappings
        fd = int(C.symm_heap_export_fd(self.h))
        got = _exchange_fds(group, {p: fd for p in range(self.world) if p != self.rank}, self.world - 1, tag + "p")
        os.close(fd)
        self.peer_va = [0] * self.world
        self.peer_va[self.rank] = self.local_va
        for src, f in got.items():
            self.peer_va[src] = int(C.symm_heap_import_peer(self.h, src, f))
            os.close(f)
        # ---- NVLS multicast
        self.mc_va = 0
        ok = torch.tensor([1 if C.symm_heap_multicast_supported(dev) else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group.pg)
        if int(ok.item()) == 1 and os.environ.get("NXDI_B200_NVLS", "1") != "0":
            if self.rank == 0:
                mfd = int(C.symm_heap_mc_create(self.h, self.world))
                _exchange_fds(group, {p: mfd for p in range(1, self.world)}, 0, tag + "m")
                os.close(mfd)
            else:
                got = _exchange_fds(group, {}, 1, tag + "m")
                C.symm_heap_mc_import(self.h, got[0])
                os.close(got[0])
            C.symm_heap_mc_add_device(self.h)
            torch.cuda.synchronize(device)
            dist.barrier(group=group.pg)          # every device has joined before anyone binds memory
            self.mc_va = int(C.symm_heap_mc_bind_map(self.h))
            torch.cuda.synchronize(device)
            dist.barrier(group=group.pg)
        self._top = _SIG_BYTES
        self.sig_ptrs = [va for va in self.peer_va]   # signal words live at offset 0 of every copy
        self._scratch_off: Optional[int] = None
        self._scratch_bytes = 0

    # ---- allocation ----------------------------------------------------------------------------------------------------
    def alloc(self, nbytes: int, align: int = 1024) -> int:
        off = (self._top + align - 1) // align * align
        if off + nbytes > self.size:
            raise MemoryError(f"symmetric heap exhausted: {off + nbytes} > {self.size}")
        self._top = off + nbytes
        return off

    def tensor(self, off: int, shape, dtype=torch.bfloat16) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        raw = self.C.symm_as_tensor(self.local_va + off, nbytes, self.device.index if self.device.index is not None else 0)
        return raw.view(dtype).view(*shape)

    def scratch(self, nbytes: int) -> int:
        """One scratch region shared by all collectives (each of them ends with a cross-GPU barrier, so the region is free when the
        next one starts); grown on demand, identically on every rank."""
        if self._scratch_off is None or nbytes > self._scratch_bytes:
            size = max(nbytes, 32 << 20)
            self._scratch_off = self.alloc(size)
            self._scratch_bytes = size
        return self._scratch_off

    @property
    def has_multicast(self) -> bool:
        return self.mc_va != 0

    # ---- collectives ---------------------------------------------------------------------------------------------------
    def _launch(self, mode: int, off: int, segs: int, rows: int, row_elems: int, residual, out):
        symm = self.group.symm
        self.C.nvls_collective(mode, self.sig_ptrs, symm.step_t, self.rank, symm.call, self.mc_va + off, self.local_va + off,
                               residual, out, segs, rows, row_elems)
        symm.call += 1
        symm.calls += 1

    def _layout(self, shape, dim: Optional[int]):
        """-> (segs, rows_per_seg, row_elems) of a contiguous tensor whose ``dim`` is the split dimension (None: flat)."""
        n = 1
        for s in shape:
            n *= int(s)
        if dim is None:
            return 1, n // 8, 8
        dim = dim % len(shape)
        segs = 1
        for s in shape[:dim]:
            segs *= int(s)
        row = 1
        for s in shape[dim + 1:]:
            row *= int(s)
        return segs, int(shape[dim]), row

    def usable(self, x: torch.Tensor, dim: Optional[int] = None, full_shape=None) -> bool:
        """Can the in-switch kernels take this tensor (``dim``: split dimension of the FULL tensor, None = flat all-reduce)?"""
        if not self.has_multicast or self.group.symm is None or not x.is_cuda or x.dtype != torch.bfloat16:
            return False
        shape = full_shape if full_shape is not None else x.shape
        segs, rows, row = self._layout(shape, dim)
        nbytes = segs * rows * row * 2
        room = self._scratch_bytes if self._scratch_off is not None and nbytes <= self._scratch_bytes else self.size - self._top - 4096
        return row % 8 == 0 and rows % self.world == 0 and rows >= self.world and 0 < nbytes <= room

    def staging(self, shape) -> torch.Tensor:
        """A tensor of ``shape`` inside the scratch region: a producer (GEMM epilogue) writes partial sums straight into it."""
        n = 1
        for s in shape:
            n *= int(s)
        return self.tensor(self.scratch(n * 2), shape)

    def in_scratch(self, x: torch.Tensor) -> bool:
        return self._scratch_off is not None and x.data_ptr() == self.local_va + self._scratch_off

    def all_reduce(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """sum over ranks (+ residual).  ``x`` may already live in the staging area (no copy-in then)."""
        x = x.contiguous()
        segs, rows, row = self._layout(x.shape, None)
        off = self.scratch(x.numel() * 2)
        if not self.in_scratch(x):
            self.tensor(off, x.shape).copy_(x)
        out = torch.empty_like(x)
        self._launch(0, off, segs, rows, row, residual.contiguous() if residual is not None else None, out)
        return out

    def reduce_scatter(self, x: torch.Tensor, dim: int, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = x.contiguous()
        segs, rows, row = self._layout(x.shape, dim)
        off = self.scratch(x.numel() * 2)
        if not self.in_scratch(x):
            self.tensor(off, x.shape).copy_(x)
        shape = list(x.shape)
        shape[dim % x.dim()] = rows // self.world
        out = torch.empty(shape, dtype=x.dtype, device=x.device)
        self._launch(1, off, segs, rows, row, residual.contiguous() if residual is not None else None, out)
        return out

    def all_gather(self, x: torch.Tensor, dim: int) -> torch.Tensor:
        """-> the gathered tensor INSIDE the staging area (valid until the next collective of this heap)."""
        x = x.contiguous()
        shape = list(x.shape)
        d = dim % x.dim()
        shape[d] *= self.world
        segs, rows, row = self._layout(shape, d)
        n = segs * rows * row
        off = self.scratch(n * 2)
        full = self.tensor(off, shape)
        full.narrow(d, self.rank * x.shape[d], x.shape[d]).copy_(x)
        self._launch(2, off, segs, rows, row, None, None)
        return full

    # ---- fused GEMM -> reduce-scatter (csrc/gemm_tcgen05.cu, rs_* parameters) -----------------------------------------------
    _RS_MAX_TILES = 16384

    def _rs_resources(self, nbytes: int):
        """Flag array [max_tiles][world] + TWO staging buffers (a rank may be one collective ahead of its slowest peer, whose
        in-switch reads of the previous staging buffer may still be in flight)."""
        if getattr(self, "_rs_flags_off", None) is None:
            self._rs_flags_off = self.alloc(self._RS_MAX_TILES * self.world * 4)
            self._rs_stage_off, self._rs_stage_bytes, self._rs_n = [None, None], 0, 0
        if nbytes > self._rs_stage_bytes:
            size = max(nbytes, 48 << 20)
            if self._top + 2 * size + 4096 > self.size:
                return False
            self._rs_stage_off = [self.alloc(size), self.alloc(size)]
            self._rs_stage_bytes = size
        return True

    def gemm_rs_usable(self, x: torch.Tensor, w: torch.Tensor, rows_per_seg: int) -> bool:
        M = x.numel() // x.shape[-1]
        return (self.has_multicast and self.group.symm is not None and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
                and w.dim() == 2 and w.is_contiguous() and w.shape[1] % 64 == 0 and w.shape[0] % 8 == 0
                and rows_per_seg % (128 * self.world) == 0 and M % rows_per_seg == 0 and M > 8
                and ((M + 127) // 128) * ((w.shape[0] + 127) // 128) <= self._RS_MAX_TILES
                and os.environ.get("NXDI_B200_FUSED_RS", "1") != "0" and self._rs_resources(M * w.shape[0] * 2))

    def gemm_reduce_scatter(self, x2d: torch.Tensor, w: torch.Tensor, bias, rows_per_seg: int, residual=None) -> torch.Tensor:
        """x2d [M, K_local] -> [M / world, N]: partial GEMM, reduce-scatter over the sequence rows and residual add in ONE kernel."""
        M, N = x2d.shape[0], w.shape[0]
        off = self._rs_stage_off[self._rs_n & 1]
        self._rs_n += 1
        staging = self.tensor(off, (M, N))
        symm = self.group.symm
        flags = [va + self._rs_flags_off for va in self.peer_va]
        out = self.C.gemm_reduce_scatter(x2d, w, bias, staging, self.mc_va + off, flags, self._RS_MAX_TILES, symm.step_t, symm.call,
                                         self.rank, rows_per_seg, residual.reshape(-1, N).contiguous() if residual is not None else None,
                                         0, [], None)
        symm.call += 1
        symm.calls += 1
        return out

    def gemm_ar_usable(self, x: torch.Tensor, w: torch.Tensor) -> bool:
        M = x.numel() // x.shape[-1]
        return self.gemm_rs_usable(x, w, M) and os.environ.get("NXDI_B200_FUSED_AR", "1") != "0"

    def gemm_all_reduce(self, x2d: torch.Tensor, w: torch.Tensor, bias, residual=None) -> torch.Tensor:
        """x2d [M, K_local] -> [M, N] = sum over ranks (+ bias on rank 0, + residual): partial GEMM, in-switch reduction of the
        tiles this rank owns, multicast of the result to every rank and the completion handshake — ONE kernel.  The result is
        copied out of the symmetric output buffer (two of them alternate; callers may keep the tensor)."""
        M, N = x2d.shape[0], w.shape[0]
        if getattr(self, "_ar_out_off", None) is None or M * N * 2 > self._ar_out_bytes:
            size = max(M * N * 2, 32 << 20)
            self._ar_out_off = [self.alloc(size), self.alloc(size)]
            self._ar_out_bytes = size
            self._ar_done_off = self.alloc(256)
            self._ar_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        off = self._rs_stage_off[self._rs_n & 1]
        out_off = self._ar_out_off[self._rs_n & 1]
        self._rs_n += 1
        staging = self.tensor(off, (M, N))
        symm = self.group.symm
        flags = [va + self._rs_flags_off for va in self.peer_va]
        done = [va + self._ar_done_off for va in self.peer_va]
        self.C.gemm_reduce_scatter(x2d, w, bias, staging, self.mc_va + off, flags, self._RS_MAX_TILES, symm.step_t, symm.call, self.rank, M,
                                   residual.reshape(-1, N).contiguous() if residual is not None else None, self.mc_va + out_off, done,
                                   self._ar_counter)
        symm.call += 1
        symm.calls += 1
        return self.tensor(out_off, (M, N)).clone()

    def close(self):
        if getattr(self, "h", None) is not None:
            self.C.symm_heap_destroy(self.h)
            self.h = None
