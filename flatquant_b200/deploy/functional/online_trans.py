"""Functional online transform + quantise (mirror of deploy/functional/online_trans.py:90-141).

`kronecker_matmul(x, invs, cmax, cmin)` returns a PackedQuantizedTensor exactly like the reference's
Triton path: `invs = [left, right]` -> L^T X R per token (kron_matmul.py), `invs = [right]` ->
per-token X R over the last axis (block_matmul.py).  One CUDA kernel: flatq_transform_quant."""
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
        xq, s16, _ = ops.transform_quant(x.contiguous(), left.contiguous(), right.contiguous(), n1, n2, fm, fn)
    elif len(invs) == 1:
        right = invs[0]
        n2 = right.shape[0]
        n1 = x.shape[-1] // n2 if x.shape[-1] != n2 else 1
        xq, s16, _ = ops.transform_quant(x.contiguous(), None, right.contiguous(), n1, n2, fm, fn)
    else:
        raise NotImplementedError
    return deploy.PackedQuantizedTensor(xq, s16.view(-1, 1))
