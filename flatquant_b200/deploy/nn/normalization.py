"""Weight-less RMSNorm with the reference's contract (deploy/nn/normalization.py:4-23): the learned scale is folded
into the following transform / linear offline, the mean runs over `mean_dim` (not x.shape[-1]) and only fp16 inputs
are promoted to fp32 for the reduction (bf16 / fp32 inputs are reduced in their own dtype, as the reference does)."""
import torch


def rms_normalize(x: torch.Tensor, mean_dim: int, eps: float) -> torch.Tensor:
    work = x.float() if x.dtype == torch.float16 else x
    inv_rms = torch.rsqrt(work.square().sum(dim=-1, keepdim=True) / mean_dim + eps)
    return (work * inv_rms).to(x.dtype)


class RMSNorm(torch.nn.Module):
    def __init__(self, mean_dim: int, eps=1e-5):
        super().__init__()
        self.mean_dim, self.eps = mean_dim, eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return rms_normalize(x, self.mean_dim, self.eps)
