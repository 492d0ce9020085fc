"""Row-parallel FlatLinear reductions over NVLink peer memory (csrc/peer_reduce.cuh, include/flatq_b200.h section (e)).

torch.distributed is the plumbing only: `torch.distributed._symmetric_memory` allocates the buffers, exchanges the
handles and maps every peer's copy (plus the NVSwitch multicast mapping) into this process; the bytes are moved by this
library's kernels -- the GEMM epilogue's peer stores (reduce-scatter) and the broadcast / in-switch all-reduce kernels.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops

_FLAG_CTAS = 96            # matches peer_grid() in csrc/api.cu


class _TpCtx(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("signal_pads", ctypes.c_void_p * 8),
                ("signal_pad_bytes", ctypes.c_size_t)]


def _ptr_array(ptrs):
    arr = (ctypes.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = int(p) if p else None
    return arr


class PeerReducer:
    """One per (capacity, width, group): serves every call with m <= capacity rows (the symmetric buffers are flat; a
    decode step and a prefill of any length share them).  `rowparallel_gemm` is the whole row-parallel linear after the
    activation quantiser: GEMM with the reduce-scatter in its epilogue + owner-side sum + broadcast (m a multiple of
    256 x world), else GEMM + the in-switch all-reduce; `all_reduce` is that all-reduce by itself."""

    def __init__(self, world, rank, dev, shape, dtype=torch.float16, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        if dtype != torch.float16:
            raise RuntimeError("PeerReducer: float16 partial products only")
        if not (2 <= world <= 8):
            raise RuntimeError("PeerReducer: 2..8 ranks")
        self.world, self.rank, self.dev, self.group = world, rank, dev, group or dist.group.WORLD
        self.M, self.N = int(shape[0]), int(shape[1])
        self.fused = self.M % (256 * world) == 0     # of a call with m == capacity (reported by the bench)
        lib = _lib.load()
        self.lib = lib

        def shared(numel, dt):
            t = symm_mem.empty(numel, dtype=dt, device=dev)
            t.zero_()
            return t, symm_mem.rendezvous(t, self.group)
        self.out, h_out = shared(self.M * self.N, torch.float16)
        self.partial, h_par = shared(self.M * self.N, torch.float16)   # staging slots (fused) or the local partial product
        self.flags, h_flg = shared(2 * _FLAG_CTAS * (world + 1), torch.int32)   # epoch flags + per-CTA counters, zeroed once
        self.out, self.partial = self.out.view(self.M, self.N), self.partial.view(self.M, self.N)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)
        self._handles = (h_out, h_par, h_flg)
        self.out_mc = int(h_out.multicast_ptr or 0)      # 0: no NVSwitch multicast on this box
        self.par_mc = int(h_par.multicast_ptr or 0) if self.out_mc else 0
        self.out_peers = _ptr_array(list(h_out.buffer_ptrs))
        self.par_peers = _ptr_array(list(h_par.buffer_ptrs))
        ctx = _TpCtx()
        ctx.rank, ctx.world = rank, world
        for r, p in enumerate(h_flg.buffer_ptrs):
            ctx.signal_pads[r] = int(p)
        ctx.signal_pad_bytes = self.flags.numel() * 4
        self.ctx = ctx
        if not self.fused and not self.out_mc:
            raise RuntimeError("PeerReducer: this shape needs the in-switch all-reduce and the box has no multicast support")
        self.name = ("GEMM epilogue peer stores (reduce-scatter) + owner sum + " +
                     "peer-store broadcast") if self.fused else \
            "in-switch all-reduce (multimem.ld_reduce / multimem.st)"

    # ---- the fused prefill path
    def rowparallel_gemm(self, xq8, s_x, w_packed, s_w, bias=None):
        """y [M, N] fp16 = sum over ranks of (xq8 . W^T) s_x s_w (+ bias once): returns this reducer's output buffer."""
        m, k = xq8.shape
        if m > self.M or w_packed.shape[0] != self.N:
            raise RuntimeError(f"PeerReducer built for up to {self.M} x {self.N}, got {(m, w_packed.shape[0])}")
        st = torch.cuda.current_stream().cuda_stream
        b = None if bias is None else bias.to(torch.float16).contiguous()
        if m % (256 * self.world) == 0:
            _lib.check(self.lib.flatq_gemm_w4a8_rowparallel(xq8.data_ptr(), s_x.data_ptr(), w_packed.data_ptr(), s_w.data_ptr(),
                                                            m, self.N, k, ctypes.addressof(self.ctx), self.par_peers, st),
                       "gemm_w4a8_rowparallel")
            _lib.check(self.lib.flatq_peer_reduce_bcast(ctypes.addressof(self.ctx), self.partial.data_ptr(), self.out_mc or None,
                                                        self.out_peers, None if b is None else b.data_ptr(), m, self.N, st),
                       "peer_reduce_bcast")
            return self.out[:m]
        part = self.partial[:m]
        ops.gemm_w4a8(xq8, s_x, w_packed, s_w, None, torch.float16, out=part)
        return self.all_reduce(part, bias=b)

    # ---- any shape: in-switch all-reduce of the symmetric partial buffer
    def partial_buffer(self):
        return self.partial

    def all_reduce(self, partial=None, bias=None):
        m = self.M if partial is None else partial.shape[0]
        if partial is not None and partial.data_ptr() != self.partial.data_ptr():
            self.partial[:m].copy_(partial)
        if not self.par_mc:
            raise RuntimeError("PeerReducer.all_reduce: no multicast support on this box")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.flatq_peer_allreduce(ctypes.addressof(self.ctx), self.par_mc, self.out_mc,
                                                 None if bias is None else bias.data_ptr(), m, self.N, st), "peer_allreduce")
        return self.out[:m]
