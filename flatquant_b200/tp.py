"""Tensor-parallel FlatLinears: the sharding the reference defines in
vllm_custom/model_executor/fake_quantized_models/llama_flatquant.py:75-126,148-289 (column-parallel qkv / gate_up,
row-parallel o / down with per-rank Kronecker factors and per-rank clip factors, `trans_list.{rank}` :551-562) and
in flatquant/model_tools/deepseekv3_utils.py:66-74,98-105 (`is_rowparallel`: F.linear -> dist.all_reduce -> bias).

  * column parallel: the input transform runs on the full (replicated) hidden state; each rank owns N/tp rows of
    the packed weight; no communication.
  * row parallel: each rank owns K/tp input features with ITS OWN Kronecker pair sized by
    get_decompose_dim(K/tp), its own per-token scale and clip factors; partial products are dequantised locally
    and summed with ONE all-reduce per linear (2 per decoder layer); the bias is added after the reduction.

One process per GPU; torch.distributed (NCCL over NVLink on B200, gloo in the CPU tests) is the plumbing.
The per-rank compute is `ops.linear_fused` (one C-ABI call); tests inject the oracle instead.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from .function_utils import get_decompose_dim


@dataclass
class Shard:
    rank: int
    tp: int
    k0: int  # input-feature range [k0, k1)
    k1: int
    n0: int  # output-feature range [n0, n1)
    n1: int
    factors: tuple  # (n1, n2) of the Kronecker pair this rank applies


def column_shard(K: int, N: int, tp: int, rank: int, align: int = 2) -> Shard:
    """Split the output features; every rank transforms the full K with the full-size factor pair."""
    if N % (tp * align):
        raise ValueError(f"N={N} is not divisible by tp*{align}={tp * align}")
    w = N // tp
    return Shard(rank, tp, 0, K, rank * w, (rank + 1) * w, get_decompose_dim(K))


def row_shard(K: int, N: int, tp: int, rank: int) -> Shard:
    """Split the input features; shard boundaries must keep 128-bit alignment of the packed rows (64 elements)."""
    if K % (tp * 64):
        raise ValueError(f"K={K} is not divisible by tp*64={tp * 64}")
    w = K // tp
    return Shard(rank, tp, rank * w, (rank + 1) * w, 0, N, get_decompose_dim(w))


def _default_linear(x, left, right, n1, n2, cmax, cmin, w_packed, s_w, bias, out_dtype):
    from . import ops
    return ops.linear_fused(x, left, right, n1, n2, cmax, cmin, w_packed, s_w, bias, out_dtype)


class ColumnParallelFlatLinear:
    """y[:, n0:n1] = FlatLinear(x) for this rank's rows of W.  `gather=True` all-gathers the column blocks."""

    def __init__(self, shard: Shard, left, right, clip, w_packed, s_w, bias=None, group=None, linear_fn=None):
        self.shard, self.left, self.right, self.clip = shard, left, right, clip
        self.w_packed, self.s_w, self.bias, self.group = w_packed, s_w, bias, group
        self.linear_fn = linear_fn or _default_linear

    def __call__(self, x, out_dtype=torch.float16, gather=False):
        n1, n2 = self.shard.factors
        y = self.linear_fn(x, self.left, self.right, n1, n2, self.clip[0], self.clip[1], self.w_packed, self.s_w,
                           self.bias, out_dtype)
        if not gather or self.shard.tp == 1:
            return y
        parts = [torch.empty_like(y) for _ in range(self.shard.tp)]
        dist.all_gather(parts, y.contiguous(), group=self.group)
        return torch.cat(parts, dim=-1)


class RowParallelFlatLinear:
    """y = all_reduce_sum_r FlatLinear_r(x[:, k0:k1]) (+ bias once, after the reduction)."""

    def __init__(self, shard: Shard, left, right, clip, w_packed, s_w, bias=None, group=None, linear_fn=None,
                 reduce_dtype=None):
        self.shard, self.left, self.right, self.clip = shard, left, right, clip
        self.w_packed, self.s_w, self.bias, self.group = w_packed, s_w, bias, group
        self.linear_fn = linear_fn or _default_linear
        self.reduce_dtype = reduce_dtype

    def __call__(self, x_local, out_dtype=torch.float16):
        n1, n2 = self.shard.factors
        y = self.linear_fn(x_local, self.left, self.right, n1, n2, self.clip[0], self.clip[1], self.w_packed, self.s_w,
                           None, out_dtype)
        if self.shard.tp > 1:
            if self.reduce_dtype is not None and y.dtype != self.reduce_dtype:
                y = y.to(self.reduce_dtype)
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)  # one SUM all-reduce per row-parallel linear
        if self.bias is not None:
            y = y + self.bias.to(y.dtype)
        return y.to(out_dtype)


def split_packed_rows(w_packed: torch.Tensor, shard: Shard) -> torch.Tensor:
    """This rank's slice of a packed [N, K/2] weight (rows for column-parallel, byte columns for row-parallel)."""
    return w_packed[shard.n0:shard.n1, shard.k0 // 2:shard.k1 // 2].contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# The reduction behind a row-parallel linear.  `make_reducer` picks the implementation once per (shape, group):
#   * PeerReducer  (flatquant_b200/peer_reduce.py): symmetric-memory buffers, the GEMM writes its partial product into
#     the local one and ONE kernel of this library reduces through the NVSwitch multicast mapping and broadcasts the sum
#     to every rank (multimem.ld_reduce / multimem.st over NVLink; csrc/peer_reduce.cuh);
#   * NcclReducer: dist.all_reduce -- the baseline form, and what the gloo CPU tests exercise.
class NcclReducer:
    name = "torch.distributed all_reduce (NCCL ring / NVLS, separate launch)"

    def __init__(self, shape, dtype, dev, group=None):
        self.buf = torch.empty(shape, dtype=dtype, device=dev)
        self.group = group

    def partial_buffer(self):
        return self.buf

    def all_reduce(self, partial, out=None):
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
        if out is not None and out.data_ptr() != partial.data_ptr():
            out.copy_(partial)
            return out
        return partial

    def rowparallel_gemm(self, xq8, s_x, w_packed, s_w, bias=None):
        from . import ops
        y = ops.gemm_w4a8(xq8, s_x, w_packed, s_w, None, self.buf.dtype, out=self.buf[:xq8.shape[0]])
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y if bias is None else y + bias.to(y.dtype)


def make_reducer(world, rank, dev, shape, dtype, group=None, prefer_peer=True):
    import os
    if prefer_peer and dev.type == "cuda" and os.environ.get("FLATQ_TP_NCCL") is None:
        try:
            from .peer_reduce import PeerReducer
            return PeerReducer(world, rank, dev, shape, dtype, group)
        except Exception as e:  # noqa: BLE001 -- no symmetric memory / multicast on this box: say so, use NCCL
            import warnings
            warnings.warn(f"flatquant_b200.tp: peer-memory reduction unavailable ({type(e).__name__}: {e}); using NCCL")
    return NcclReducer(shape, dtype, dev, group)
