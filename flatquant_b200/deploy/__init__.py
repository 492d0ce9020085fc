"""Real-quant operator API (mirror of the reference's `deploy` package, deploy/__init__.py:37-71):
`matmul`, `sym_quant`, `sym_dequant`, `PackedQuantizedTensor`, sub-packages `nn` and `functional`.

The reference routes these to its pybind module `deploy._CUDA` (deploy/kernels/bindings.cpp);
here they call the C ABI of libflatq_b200.so.  Assertions / error behaviour follow the reference
(bare `assert` on dtype/shape, RuntimeError from the native layer)."""
import torch

from .. import ops


def _flat2d(x):
    return x.reshape(-1, x.shape[-1]), x.shape[:-1]


class PackedQuantizedTensor:
    """A 4-bit packed activation and its per-token scales (deploy/__init__.py:55-71)."""

    def __init__(self, quantized_x: torch.Tensor, scales_x: torch.Tensor):
        self.quantized_x = quantized_x
        self.scales_x = scales_x

    def size(self):
        return self.quantized_x.size()

    @property
    def device(self):
        return self.quantized_x.device

    @property
    def dtype(self):
        return self.quantized_x.dtype


def matmul(A, B):
    """int32 [..., N] = A[..., K/2] (packed) . B[N, K/2]^T (packed)   (deploy/__init__.py:37-41)."""
    assert A.shape[-1] % 32 == 0, "A.shape[-1]: {} must be multiplication of 32".format(A.shape[-1])
    a2, a_lead = _flat2d(A)
    b2, b_lead = _flat2d(B)
    return ops.gemm_i4(a2.contiguous(), b2.contiguous()).view(*a_lead, *b_lead)


def sym_quant(x, scale):
    """packed u8 [..., K/2] = clamp(rn(x / scale[row]), -8, 7)   (deploy/__init__.py:43-46)."""
    assert x.dtype == scale.dtype == torch.float16
    x2, lead = _flat2d(x)
    return ops.sym_quant(x2.contiguous(), scale.reshape(-1).contiguous()).view(*lead, -1)


def sym_dequant(q, scale_row, scale_col, bits=32):
    """fp16 [..., N] = q * scale_row[m] * scale_col[n]   (deploy/__init__.py:48-52)."""
    assert q.dtype == torch.int32
    assert scale_row.dtype == scale_col.dtype == torch.float16
    if bits != 32:
        raise RuntimeError("Unsupported data type")  # bindings.cpp:82-83
    q2, lead = _flat2d(q)
    return ops.sym_dequant(q2.contiguous(), scale_row.reshape(-1).contiguous(),
                           scale_col.reshape(-1).contiguous()).view(*lead, -1)


from . import functional, nn  # noqa: E402,F401

__all__ = ["matmul", "sym_quant", "sym_dequant", "PackedQuantizedTensor", "nn", "functional"]
