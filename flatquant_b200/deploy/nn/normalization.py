"""Weight-less RMSNorm (mirror of deploy/nn/normalization.py): the learned scale is folded into the
following transform / linear offline."""
import torch


class RMSNorm(torch.nn.Module):
    def __init__(self, mean_dim: int, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.mean_dim = mean_dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dt = x.dtype
        xf = x.float()
        var = xf.pow(2).sum(-1, keepdim=True) / self.mean_dim
        return (xf * torch.rsqrt(var + self.eps)).to(dt)
