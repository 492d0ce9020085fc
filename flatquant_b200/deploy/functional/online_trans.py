"""Functional online transform + quantise (mirror of deploy/functional/online_trans.py:90-141).

`kronecker_matmul(x, invs, cmax, cmin)` returns a PackedQuantizedTensor exactly like the reference's
Triton path: `invs = [left, right]` -> L^T X R per token (kron_matmul.py), `invs = [inv]` with a 4-D
[bsz, seq, head_dim, heads] input -> block_matmul (one scale per token, heads-major packing).  One CUDA kernel:
flatq_transform_quant."""
import math

import torch

from ... import ops
from ...function_utils import get_decompose_dim  # noqa: F401  (re-exported like the reference)
from .hadamard import get_hadK, is_pow2, matmul_hadU_cuda  # noqa: F401  (the reference defines them in this module)


def quant(x, clip_factor_a_max=1.0, clip_factor_a_min=1.0, input_clip_ratio=1.0):
    """Per-token absmax quantise + pack of an fp16 tensor (online_trans.py:90-110)."""
    import flatquant_b200.deploy as deploy
    cmax, cmin = float(clip_factor_a_max), float(clip_factor_a_min)
    if cmax != 1.0:
        fm, fn = 1 / (1 + math.exp(-cmax)), 1 / (1 + math.exp(-cmin))
    else:
        fm = fn = float(input_clip_ratio)
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    xq, s16, _ = ops.quantize_rows(x2, fm, fn, half_math=True)
    return deploy.PackedQuantizedTensor(xq.view(*x.shape[:-1], -1), s16.view(-1, 1))


def kronecker_matmul(x, invs, clip_factor_a_max=1.0, clip_factor_a_min=1.0):
    import flatquant_b200.deploy as deploy
    # kron_matmul.py:86-92 applies sigmoid to whatever factor it is given
    fm = 1 / (1 + math.exp(-float(clip_factor_a_max)))
    fn = 1 / (1 + math.exp(-float(clip_factor_a_min)))
    if len(invs) == 2:
        assert x.dim() == 3, "kronecker_matmul expects [bsz, seq, hidden]"
        left, right = invs
        n1, n2 = left.shape[0], right.shape[0]
        bsz, seq_len, _ = x.shape
        xq, s16, _ = ops.transform_quant(x.contiguous(), left.contiguous(), right.contiguous(), n1, n2, fm, fn)
        # online_trans.py:121-122
        return deploy.PackedQuantizedTensor(xq.reshape(bsz, seq_len, -1), s16.reshape(bsz, 1, seq_len))
    elif len(invs) == 1:
        # online_trans.py:124-139: the o_proj transform over the heads axis.  The caller hands over
        # [bsz, seq, head_dim, heads]; ONE scale per token over the whole head_dim x heads block, codes packed in
        # the transposed (heads-major) order by block_matmul, quantized_x reshaped to (bsz, seq, -1, heads).
        assert x.dim() == 4, "single-matrix kronecker_matmul expects [bsz, seq, head_dim, num_heads]"
        from ..kernels.block_matmul import block_matmul
        bsz, seq_len, head_dim, num_heads = x.shape
        packed = block_matmul(x.reshape(-1, head_dim, num_heads).contiguous(), invs[0].contiguous(), seq_len,
                              clip_factor_a_max, clip_factor_a_min)
        packed.quantized_x = packed.quantized_x.reshape(bsz, seq_len, -1, num_heads)
        packed.scales_x = packed.scales_x.reshape(bsz, 1, seq_len)
        return packed
    else:
        raise NotImplementedError
