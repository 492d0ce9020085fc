"""`kron_matmul(a, b, c, seq_len, cmax, cmin)` (mirror of deploy/kernels/kron_matmul.py:192-266, a Triton kernel pair in
the reference): a = L^T [n1, n1] (the caller passes the transposed left factor, online_trans.py:118), b = tokens
[B, n1, n2], c = R [n2, n2]; returns PackedQuantizedTensor(quantized [B, n1*n2/2] u8, scales [B, 1] f16).
One sm_100a kernel (flatq_transform_quant); the raw clip factors go through a sigmoid as in kron_matmul.py:86-92."""
import math

from ... import ops


def kron_matmul(a, b, c, seq_len, clip_factor_a_max, clip_factor_a_min):
    import flatquant_b200.deploy as deploy
    assert a.shape[1] == b.shape[1], "Incompatible dimensions"
    assert b.shape[2] == c.shape[0], "Incompatible dimensions"
    assert a.is_contiguous() and b.is_contiguous() and c.is_contiguous(), "Matrices must be contiguous"
    B, n1, n2 = b.shape
    fm = 1 / (1 + math.exp(-float(clip_factor_a_max)))
    fn = 1 / (1 + math.exp(-float(clip_factor_a_min)))
    left = a.t().contiguous()  # the kernel applies L^T itself
    xq, s16, _ = ops.transform_quant(b.reshape(B, n1 * n2), left, c, n1, n2, fm, fn)
    return deploy.PackedQuantizedTensor(xq.reshape(B, -1), s16.view(B, 1))
