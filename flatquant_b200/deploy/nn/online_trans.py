"""OnlineTrans (mirror of deploy/nn/online_trans.py:18-67): online Kronecker ("matmul") transform
fused with per-token INT4 quantisation; buffers `left_matrix`, `right_matrix`, `diag_scale`,
`clip_factor_a_max/min` keep the reference names so checkpoints load unchanged."""
import torch

from ..functional import online_trans as F_ot
from ...function_utils import get_decompose_dim  # noqa: F401


def _placeholder_buffers(trans, trans_dim, decompose):
    """name -> shape of the fp16 buffers a checkpoint fills (uninitialised placeholders until then)."""
    if trans != "matmul":
        return {}
    if not decompose:
        return {"right_matrix": (trans_dim, trans_dim)}
    n1, n2 = get_decompose_dim(trans_dim)
    return {"left_matrix": (n1, n1), "right_matrix": (n2, n2), "diag_scale": (trans_dim,)}


class OnlineTrans(torch.nn.Module):
    def __init__(self, trans_dim, force_fp32=False, trans="had", decompose=True, lac=False):
        super().__init__()
        self.trans_dim, self.trans, self.decompose = trans_dim, trans, decompose
        self.fp32_trans, self.lac = force_fp32, lac
        for name, shape in _placeholder_buffers(trans, trans_dim, decompose).items():
            self.register_buffer(name, torch.randn(shape, dtype=torch.float16))
        if trans == "had":
            from ..functional.hadamard import get_hadK
            rem, self.rem_dim = get_hadK(trans_dim)
            if rem is None:
                self.had_rem_dim = None
            else:
                self.register_buffer("had_rem_dim", rem if force_fp32 else rem.to(torch.float16))
        for name in ("clip_factor_a_max", "clip_factor_a_min"):  # raw LAC factors; 1.0 = "not calibrated"
            self.register_buffer(name, torch.tensor(1.0))

    def _factors(self):
        return [getattr(self, n) for n in ("left_matrix", "right_matrix") if hasattr(self, n)]

    def forward(self, x):
        x = x.float() if self.fp32_trans else x
        if self.trans == "matmul":
            return F_ot.kronecker_matmul(x, self._factors(), float(self.clip_factor_a_max), float(self.clip_factor_a_min))
        if self.trans == "had":
            from ..functional.hadamard import matmul_hadU_cuda
            return matmul_hadU_cuda(x, self.had_rem_dim, self.rem_dim)
        return x
