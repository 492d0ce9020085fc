"""Online Hadamard transform (mirror of the Hadamard part of deploy/functional/online_trans.py:21-71,144-151).

`matmul_hadU_cuda(X, hadK, K)` = X viewed as [..., K, n/K]: hadamard_transform over the last axis (scaled by 1/sqrt(n),
fp16), then `hadK @` over the K axis.  Power-of-two sizes run on the sm_100a FHT kernel (flatq_hadamard, replacing
third-party/fast-hadamard-transform); n = K * 2^m runs on ONE fused kernel (flatq_hadamard_kn: FHT segments + the dense
K x K factor with the row resident in shared memory) where the reference launches the FHT and a cuBLAS batched GEMM.

`get_hadK(n)` picks K like the reference (the first of 172, 156, 140, 108, 60, 52, 36, 28, 40, 20, 12 that divides n
with a power-of-two cofactor) and returns the reference's own K x K matrix: the tables of online_trans.py:161-4083
(Sloane's library of Hadamard matrices) are shipped as DATA, one bit per entry, in data/hadamard_tables.npz, written
from the reference by tools/make_hadamard_tables.py and checked entry for entry against it in
tests/test_dropin_reference_wrappers.py.  (The reference's exporter writes parameters only, so `had_rem_dim` never
travels in a checkpoint: the online factor MUST be the table the weights were rotated with offline.)  `paley_hadamard`
-- a number-theoretic construction of a Hadamard matrix of the same orders -- is kept as an independent check that the
shipped tables are Hadamard matrices of the right order; it is not used on the inference path."""
import math
import os

import torch

from ... import ops

_ORDERS = (172, 156, 140, 108, 60, 52, 36, 28, 40, 20, 12)  # the reference's search order (online_trans.py:21-71)


def is_pow2(n):
    return n > 0 and (n & (n - 1)) == 0


def _is_prime(q):
    return q > 1 and all(q % d for d in range(2, int(q ** 0.5) + 1))


def _jacobsthal(q):
    """Q[i, j] = chi(i - j) over GF(q), q prime: +1 for non-zero squares, -1 for non-squares, 0 on the diagonal."""
    squares = {(x * x) % q for x in range(1, q)}
    idx = torch.arange(q)
    diff = (idx[:, None] - idx[None, :]) % q
    chi = torch.tensor([0] + [1 if r in squares else -1 for r in range(1, q)], dtype=torch.float32)
    return chi[diff]


def paley_hadamard(order):
    """A Hadamard matrix of the given order by Paley's construction I (order = q + 1, q prime, q = 3 mod 4) or II
    (order = 2 (q + 1), q prime, q = 1 mod 4), or by Sylvester doubling of one of those; None when neither applies."""
    q = order - 1
    if _is_prime(q) and q % 4 == 3:
        Q = _jacobsthal(q)
        S = torch.zeros(order, order)
        S[0, 1:], S[1:, 0], S[1:, 1:] = 1.0, -1.0, Q
        return S + torch.eye(order)
    if order % 2 == 0:
        q = order // 2 - 1
        if _is_prime(q) and q % 4 == 1:
            Q = _jacobsthal(q)
            n = q + 1
            S = torch.zeros(n, n)
            S[0, 1:], S[1:, 0], S[1:, 1:] = 1.0, 1.0, Q
            one, eye = torch.tensor([[1.0, 1.0], [1.0, -1.0]]), torch.tensor([[1.0, -1.0], [-1.0, -1.0]])
            return torch.kron(S, one) + torch.kron(torch.eye(n), eye)
        half = paley_hadamard(order // 2)
        if half is not None:
            return torch.kron(torch.tensor([[1.0, 1.0], [1.0, -1.0]]), half)
    return None


def hadK_order(n):
    """K of get_hadK: 1 for a power of two, else the first reference order dividing n with a power-of-two cofactor."""
    if is_pow2(n):
        return 1
    for k in _ORDERS:
        if n % k == 0 and is_pow2(n // k):
            return k
    raise NotImplementedError(f"get_hadK({n}): no Hadamard factorisation K x 2^m with K in {_ORDERS}")


_TABLES = {}


def hadamard_table(k):
    """The reference's K x K Hadamard matrix (entries +-1, fp32) from the shipped bit table."""
    if k not in _TABLES:
        import numpy as np
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hadamard_tables.npz")
        with np.load(path) as z:
            bits = np.unpackbits(z[f"had{k}"])[: k * k].reshape(k, k)
        _TABLES[k] = torch.from_numpy(bits.astype(np.float32) * 2.0 - 1.0)
    return _TABLES[k].clone()


def get_hadK(n, transpose=False):
    """(hadK [K, K] fp32 or None, K) -- online_trans.py:21-71."""
    k = hadK_order(n)
    if k == 1:
        return None, 1
    h = hadamard_table(k)
    return (h.t().contiguous() if transpose else h), k


def matmul_hadU_cuda(X, hadK, K):
    n = X.shape[-1]
    if K == 1:
        return ops.hadamard(X.contiguous(), 1.0 / math.sqrt(n))
    if hadK is None:
        raise RuntimeError(f"matmul_hadU_cuda: K = {K} needs its K x K Hadamard factor")
    return ops.hadamard_kn(X.contiguous(), hadK, K, 1.0 / math.sqrt(n))
