"""OnlineTrans (mirror of deploy/nn/online_trans.py:18-67): online Kronecker ("matmul") transform
fused with per-token INT4 quantisation; buffers `left_matrix`, `right_matrix`, `diag_scale`,
`clip_factor_a_max/min` keep the reference names so checkpoints load unchanged."""
import torch

from ..functional import online_trans as F_ot
from ...function_utils import get_decompose_dim  # noqa: F401


class OnlineTrans(torch.nn.Module):
    def __init__(self, trans_dim, force_fp32=False, trans="had", decompose=True, lac=False):
        super().__init__()
        self.fp32_trans = force_fp32
        self.trans = trans
        self.decompose = decompose
        self.trans_dim = trans_dim
        if trans == "had":
            from ..functional.hadamard import get_hadK
            had_rem, self.rem_dim = get_hadK(trans_dim)
            if had_rem is not None:
                self.register_buffer("had_rem_dim", had_rem if force_fp32 else had_rem.to(torch.float16))
            else:
                self.had_rem_dim = None
        elif trans == "matmul":
            if decompose:
                n1, n2 = get_decompose_dim(trans_dim)
                self.register_buffer("left_matrix", torch.randn([n1, n1], dtype=torch.float16))
                self.register_buffer("right_matrix", torch.randn([n2, n2], dtype=torch.float16))
                self.register_buffer("diag_scale", torch.randn([trans_dim], dtype=torch.float16))
            else:
                self.register_buffer("right_matrix", torch.randn([trans_dim, trans_dim], dtype=torch.float16))
        self.lac = lac
        self.register_buffer("clip_factor_a_max", torch.tensor(1.0))
        self.register_buffer("clip_factor_a_min", torch.tensor(1.0))

    def forward(self, x):
        if self.fp32_trans:
            x = x.float()
        if self.trans == "had":
            from ..functional.hadamard import matmul_hadU_cuda
            return matmul_hadU_cuda(x, self.had_rem_dim, self.rem_dim)
        if self.trans == "matmul":
            invs = [getattr(self, n) for n in ("left_matrix", "right_matrix") if hasattr(self, n)]
            return F_ot.kronecker_matmul(x, invs, float(self.clip_factor_a_max), float(self.clip_factor_a_min))
        return x
