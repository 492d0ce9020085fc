"""Shape contract + small linear-algebra helpers (mirror of flatquant/function_utils.py)."""
import math

import numpy as np
import torch


def get_decompose_dim(n):
    """Kronecker factor sizes (n1, n2), n1*n2 == n, n1 <= n2 closest to sqrt(n).

    Same contract as flatquant/function_utils.py:11-21 (Fermat-style search: the smallest
    a >= ceil(sqrt(n)) with a^2 - n a perfect square b^2 gives (a-b, a+b))."""
    a = math.isqrt(n)
    if a * a != n:
        a += 1
    b = math.isqrt(a * a - n)
    while b * b != a * a - n:
        a += 1
        b = math.isqrt(a * a - n)
    return a - b, a + b


def get_init_scale(w_smax, x_smax, alpha=0.5):
    """SmoothQuant-style diagonal init (flatquant/function_utils.py:7-8)."""
    return (w_smax.pow(1 - alpha) / x_smax.pow(alpha)).clamp(min=1e-5)


def get_random_orthg(size):
    """Haar-random orthogonal matrix from the numpy RNG, fp64 (flatquant/function_utils.py:24-28)."""
    q, r = np.linalg.qr(np.random.randn(size, size))
    return torch.from_numpy(q * np.sign(np.diag(r))[None, :])


def get_init_weight(dim):
    return get_random_orthg(dim)


def get_inverse(matrix):
    """fp64 inverse cast back to the input dtype (flatquant/function_utils.py:35-37)."""
    return torch.linalg.inv(matrix.double()).to(matrix.dtype)


def get_paras_dict_by_name(model, required_names, destination=None, prefix=""):
    """Parameters whose name contains one of `required_names`, grouped in the order of `required_names`
    (flatquant/function_utils.py:51-58) -- the layout of flat_matrices.pth."""
    from collections import OrderedDict
    out = OrderedDict() if destination is None else destination
    for wanted in required_names:
        for name, param in model.named_parameters():
            if wanted in name:
                out[prefix + name] = param.detach()
    return out
