"""Affine-transform modules (mirror of flatquant/trans_utils.py): same class names, constructor
signatures, attribute / buffer names (state-dict keys are part of the checkpoint contract:
flat_utils.py:65-78, deploy/transformers/modeling_llama.py:479-514) and `forward(inp, inv_t=False)`,
`get_matrix(inv_t)`, `to_eval_mode()`.

Eval-mode forward on CUDA fp16 activations runs the sm_100a kernel via kronecker_matmul /
ops.kron_transform; learnable (train-mode) forms are torch autograd code used by calibration only."""
import torch
import torch.nn as nn

from . import ops
from .flat_utils import kronecker_matmul
from .function_utils import get_init_weight, get_inverse


def _cayley_linear(size):
    """nn.Linear(size,size) initialised to a random orthogonal matrix and constrained to stay
    orthogonal through the Cayley map (trans_utils.py:11-16)."""
    lin = nn.Linear(size, size, bias=False, dtype=torch.float32)
    lin.weight.data = get_init_weight(size).to(lin.weight)
    return nn.utils.parametrizations.orthogonal(lin, orthogonal_map="cayley", use_trivialization=False)


def _usv(u, diag, v, inverse=False):
    d = 1 / diag if inverse else diag
    return u.weight @ torch.diag(d) @ v.weight.t()


def _matmul_last(inp, mat):
    """[..., n] @ mat for a single (non-decomposed) transform.  16-bit inference activations on CUDA with n <= 256 (head
    and head-count sized transforms: o_trans, the K / V cache transforms) run on the sm_100a kernel; a larger single
    matrix is a plain dense GEMM, which is what a library GEMM is for; fp32 / autograd / CPU follow the reference."""
    from .flat_utils import _inference_act, as_kernel_f16
    n = mat.shape[0]
    if _inference_act(inp) and n <= 256:
        if n % 16:
            raise RuntimeError(f"single transform of size {n}: no sm_100a kernel (n % 16 != 0); use fp32 activations for the "
                               "reference's torch path")
        flat = as_kernel_f16(inp.reshape(-1, n))
        return ops.kron_transform(flat, None, mat.to(inp).to(flat).contiguous(), 1, n).view(inp.shape).to(inp.dtype)
    return inp.reshape(-1, n).matmul(mat).reshape(inp.shape)


class _SingleBase(nn.Module):
    _eval_mode = False

    def forward(self, inp, inv_t=False):
        return _matmul_last(inp, self.get_matrix(inv_t=inv_t).to(inp))

    def _frozen(self, inv_t):
        return self.matrix_inv_t if inv_t else self.matrix

    def _freeze(self, matrix, matrix_inv_t):
        self.matrix = nn.Parameter(matrix, requires_grad=False)
        self.matrix_inv_t = nn.Parameter(matrix_inv_t, requires_grad=False)
        self._eval_mode = True

    def __repr__(self):
        shape = self.matrix.shape if hasattr(self, "matrix") else self._train_shape()
        return f"{type(self).__name__}(eval_mode={self._eval_mode}, matrix.shape={shape})"


class SVDSingleTransMatrix(_SingleBase):
    """P = U diag(s) V^T with Cayley-orthogonal U, V (trans_utils.py:8-54)."""

    def __init__(self, size):
        super().__init__()
        self.linear_u = _cayley_linear(size)
        self.linear_v = _cayley_linear(size)
        self.linear_diag = nn.Parameter(torch.ones(size, dtype=torch.float32))
        self._eval_mode = False

    def _train_shape(self):
        return self.linear_u.weight.shape

    def get_matrix(self, inv_t=False):
        if self._eval_mode:
            return self._frozen(inv_t)
        return _usv(self.linear_u, self.linear_diag, self.linear_v, inverse=inv_t)

    def to_eval_mode(self):
        if not self._eval_mode:
            m = _usv(self.linear_u, self.linear_diag, self.linear_v)
            mi = _usv(self.linear_u, self.linear_diag, self.linear_v, inverse=True)
            self._freeze(m, mi)
            del self.linear_u, self.linear_diag, self.linear_v


class InvSingleTransMatrix(_SingleBase):
    """Plain learnable matrix with an fp64 inverse (trans_utils.py:128-170)."""

    def __init__(self, size):
        super().__init__()
        self.linear = nn.Linear(size, size, bias=False)
        self.linear.weight.data = get_init_weight(size).to(self.linear.weight)
        self._eval_mode = False

    def _train_shape(self):
        return self.linear.weight.shape

    def get_matrix(self, inv_t=False):
        if self._eval_mode:
            return self._frozen(inv_t)
        return get_inverse(self.linear.weight).T if inv_t else self.linear.weight

    def to_eval_mode(self):
        if not self._eval_mode:
            self._freeze(self.linear.weight, get_inverse(self.linear.weight).T)


class _DecomposeBase(nn.Module):
    """Shared logic of the Kronecker-decomposed transforms: optional per-channel `diag_scale`
    (folded into the preceding RMSNorm at inference: flat_utils.py:20-27 sets use_diag=False),
    then kronecker_matmul with (matrix_left, matrix_right) or their inverse-transposes."""

    def _init_diag(self, left_size, right_size, add_diag, diag_init_para, dtype=None):
        self.add_diag, self.use_diag = add_diag, True
        if add_diag:
            init = torch.ones(left_size * right_size, dtype=dtype) if diag_init_para is None else diag_init_para
            self.diag_scale = nn.Parameter(init, requires_grad=True)
        self._eval_mode = False

    def forward(self, inp, inv_t=False):
        if self.add_diag and self.use_diag:
            d = self.diag_scale.to(inp)
            inp = inp / d if inv_t else inp * d
        left, right = self._factors(inv_t)
        return kronecker_matmul(inp, left.to(inp), right.to(inp))

    def _factors(self, inv_t):
        if self._eval_mode:
            return (self.matrix_left_inv, self.matrix_right_inv) if inv_t else (self.matrix_left, self.matrix_right)
        return self._train_factors(inv_t)

    def _freeze(self, left, right, left_inv, right_inv):
        for name, val in (("matrix_left", left), ("matrix_right", right), ("matrix_left_inv", left_inv),
                          ("matrix_right_inv", right_inv)):
            setattr(self, name, nn.Parameter(val, requires_grad=False))
        self._eval_mode = True

    def __repr__(self):
        if hasattr(self, "matrix_left"):
            return (f"{type(self).__name__}(_eval_mode={self._eval_mode}, matrix.shape={self.matrix_left.shape}, "
                    f"matrix_right.shape={self.matrix_right.shape}, )")
        return f"{type(self).__name__}(_eval_mode={self._eval_mode})"


class SVDDecomposeTransMatrix(_DecomposeBase):
    """P = P_left (x) P_right, each factor U diag(s) V^T (trans_utils.py:57-124)."""

    def __init__(self, left_size, right_size, add_diag=False, diag_init_para=None):
        super().__init__()
        self.linear_u_left, self.linear_v_left = _cayley_linear(left_size), _cayley_linear(left_size)
        self.linear_diag_left = nn.Parameter(torch.ones(left_size, dtype=torch.float32))
        self.linear_u_right, self.linear_v_right = _cayley_linear(right_size), _cayley_linear(right_size)
        self.linear_diag_right = nn.Parameter(torch.ones(right_size, dtype=torch.float32))
        self._init_diag(left_size, right_size, add_diag, diag_init_para, dtype=torch.float32)

    def _train_factors(self, inv_t):
        return (_usv(self.linear_u_left, self.linear_diag_left, self.linear_v_left, inverse=inv_t),
                _usv(self.linear_u_right, self.linear_diag_right, self.linear_v_right, inverse=inv_t))

    def to_eval_mode(self):
        if not self._eval_mode:
            left, right = self._train_factors(False)
            left_inv, right_inv = self._train_factors(True)
            self._freeze(left, right, left_inv, right_inv)
            del self.linear_u_left, self.linear_diag_left, self.linear_v_left
            del self.linear_u_right, self.linear_diag_right, self.linear_v_right


class InvDecomposeTransMatrix(_DecomposeBase):
    """Kronecker factors as plain matrices with fp64 inverses (trans_utils.py:173-221)."""

    def __init__(self, left_size, right_size, add_diag=False, diag_init_para=None):
        super().__init__()
        self.linear_left = nn.Linear(left_size, left_size, bias=False)
        self.linear_left.weight.data = get_init_weight(left_size).to(self.linear_left.weight)
        self.linear_right = nn.Linear(right_size, right_size, bias=False)
        self.linear_right.weight.data = get_init_weight(right_size).to(self.linear_right.weight)
        self._init_diag(left_size, right_size, add_diag, diag_init_para)

    def _train_factors(self, inv_t):
        if inv_t:
            return get_inverse(self.linear_left.weight).T, get_inverse(self.linear_right.weight).T
        return self.linear_left.weight, self.linear_right.weight

    def to_eval_mode(self):
        if not self._eval_mode:
            left_inv, right_inv = self._train_factors(True)
            self._freeze(self.linear_left.weight, self.linear_right.weight, left_inv, right_inv)
            del self.linear_left, self.linear_right
