"""kronecker_matmul -- the online transform (mirror of flatquant/flat_utils.py:6-17)."""
import torch

from . import ops

_KERNEL_MAX = 256


def _kernel_ok(x, n1, n2):
    return (x.is_cuda and x.dtype == torch.float16 and n2 % 16 == 0 and n1 <= _KERNEL_MAX and n2 <= _KERNEL_MAX
            and (n1 + 15) // 16 * 32 <= (512 if n2 <= 128 else 256) and not torch.is_grad_enabled())


def kronecker_matmul(x, hadL, hadR):
    """x @ kron(hadL, hadR), evaluated per token as hadL^T . reshape(x, n1, n2) . hadR.

    Inference activations (CUDA, fp16, no grad) run on the sm_100a kernel (flatq_kron_transform).
    Everything else -- the offline fp64 weight reparameterisation and calibration-time autograd,
    which are not on the inference hot path -- is expressed with torch.matmul as in the reference."""
    n1, n2 = hadL.shape[0], hadR.shape[0]
    if _kernel_ok(x, n1, n2):
        xc = x.contiguous()
        return ops.kron_transform(xc, hadL.to(xc).contiguous(), hadR.to(xc).contiguous(), n1, n2).view(x.shape)
    lead = x.shape
    t = x.reshape(-1, n1, n2)
    return (hadL.transpose(0, 1) @ (t @ hadR)).reshape(lead)
