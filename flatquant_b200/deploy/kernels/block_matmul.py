"""`block_matmul(b, c, seq_len, cmax, cmin, just_quantize=False)` (mirror of deploy/kernels/block_matmul.py:231-311):
per token b_t [M, N] @ c [N, N] (the single-matrix o_proj transform over the heads axis), quantised and packed in
the TRANSPOSED order ([N, M] row-major per token, block_matmul.py:86) that o_proj's weight layout expects."""
import math

import torch

from ... import ops


def block_matmul(b, c, seq_len, clip_factor_a_max, clip_factor_a_min, just_quantize=False):
    import flatquant_b200.deploy as deploy
    B, M, N = b.shape
    fm = 1 / (1 + math.exp(-float(clip_factor_a_max)))
    fn = 1 / (1 + math.exp(-float(clip_factor_a_min)))
    # (b_t @ c)^T = c^T @ b_t^T : a left-only transform of the transposed token with L = c
    bt = b.transpose(1, 2).contiguous().reshape(B, N * M)
    if just_quantize:
        xq, s16, _ = ops.quantize_rows(bt, fm, fn, half_math=False)
    elif M % 16 == 0:
        xq, s16, _ = ops.transform_quant(bt, c.contiguous(), None, N, M, fm, fn)
    else:
        z = torch.matmul(c.t().to(bt.dtype), bt.view(B, N, M)).reshape(B, N * M)  # ragged head_dim: not a deployed shape
        xq, s16, _ = ops.quantize_rows(z.contiguous(), fm, fn, half_math=False)
    return deploy.PackedQuantizedTensor(xq.reshape(B, -1), s16.view(B, 1))
