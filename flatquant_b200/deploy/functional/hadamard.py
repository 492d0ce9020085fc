"""Online Hadamard transform (mirror of the Hadamard part of deploy/functional/online_trans.py:21-71,144-151).

`matmul_hadU_cuda(X, hadK, K)` = X . (H_K (x) H_{n/K}) / sqrt(n).  The power-of-two factor runs on the sm_100a FHT
kernel (flatq_hadamard, replacing third-party/fast-hadamard-transform); a non-trivial K x K factor `hadK` is applied
as a small left transform.  The reference hard-codes its 12/20/28/36/40/... Hadamard matrices in a 4000-line table;
they are data, not code, and are not reproduced here: `get_hadK` returns (None, 1) for powers of two and asks the
caller to supply the matrix (e.g. from a checkpoint's `had_rem_dim` buffer) otherwise."""
import math

import torch

from ... import ops


def is_pow2(n):
    return n > 0 and (n & (n - 1)) == 0


def get_hadK(n, transpose=False):
    if is_pow2(n):
        return None, 1
    raise NotImplementedError(
        f"get_hadK({n}): the size is not a power of two; pass the K x K Hadamard factor explicitly "
        "(OnlineTrans.had_rem_dim buffer) -- the reference's hard-coded tables are not bundled")


def matmul_hadU_cuda(X, hadK, K):
    n = X.shape[-1]
    if K == 1:
        return ops.hadamard(X.contiguous(), 1.0 / math.sqrt(n))
    # X viewed as [..., K, n/K]: FHT over the last axis, then the dense K x K factor over the K axis
    inp = ops.hadamard(X.contiguous().view(-1, K, n // K), 1.0)
    out = torch.matmul(hadK.to(inp).unsqueeze(0), inp)  # small dense factor (K <= 172)
    return (out / math.sqrt(n)).reshape(X.shape)
