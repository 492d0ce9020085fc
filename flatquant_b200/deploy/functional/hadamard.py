"""Online Hadamard transform (mirror of the Hadamard part of deploy/functional/online_trans.py:21-71,144-151).

`matmul_hadU_cuda(X, hadK, K)` = X . (H_K (x) H_{n/K}) / sqrt(n).  The power-of-two factor runs on the sm_100a FHT
kernel (flatq_hadamard, replacing third-party/fast-hadamard-transform); a non-trivial K x K factor `hadK` is applied
as a small dense transform over the K axis.

`get_hadK(n)` picks K like the reference (the first of 172, 156, 140, 108, 60, 52, 36, 28, 40, 20, 12 that divides n
with a power-of-two cofactor).  The reference hard-codes its K x K matrices as a 4000-line table of Sloane's library;
that is data and is not reproduced here.  Where number theory gives a matrix of the same order -- Paley's constructions
over a prime field: 12, 20, 28, 36, 60, 108, 140, and 40 by doubling -- `get_hadK` BUILDS a Hadamard matrix of that
order.  It is a valid H (H.H^T = K.I) but not entry-for-entry the reference's: a checkpoint whose weights were rotated
offline with the reference's table carries that table in its `had_rem_dim` buffer (OnlineTrans registers it, so it is
in the state dict), and loading the checkpoint overwrites the generated one.  Orders without a prime-field Paley
construction (52, 156, 172) get a placeholder that must come from a checkpoint."""
import math

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


def get_hadK(n, transpose=False):
    """(hadK [K, K] fp32 or None, K).  For orders without a construction the matrix is a NaN placeholder: it must be
    overwritten from a checkpoint (`OnlineTrans.had_rem_dim`); OnlineTrans refuses to run with it."""
    k = hadK_order(n)
    if k == 1:
        return None, 1
    h = paley_hadamard(k)
    if h is None:
        h = torch.full((k, k), float("nan"))
    return (h.t().contiguous() if transpose else h), k


def matmul_hadU_cuda(X, hadK, K):
    n = X.shape[-1]
    if K == 1:
        return ops.hadamard(X.contiguous(), 1.0 / math.sqrt(n))
    if hadK is None:
        raise RuntimeError(f"matmul_hadU_cuda: K = {K} needs its K x K Hadamard factor")
    # X viewed as [..., K, n/K]: FHT over the last axis, then the dense K x K factor over the K axis
    inp = ops.hadamard(X.contiguous().view(-1, K, n // K), 1.0)
    out = torch.matmul(hadK.to(inp).unsqueeze(0), inp)  # small dense factor (K <= 172)
    return (out / math.sqrt(n)).reshape(X.shape)
