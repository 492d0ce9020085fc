"""Quantizer (mirror of deploy/nn/quantization.py:5-36): per-token symmetric INT4 activation
quantiser; passes an already packed input through."""
import torch

from ... import ops


class Quantizer(torch.nn.Module):
    def __init__(self, input_clip_ratio=1.0, lac=False):
        super().__init__()
        self.input_clip_ratio = input_clip_ratio
        self.lac = lac
        self.register_buffer("clip_factor_a_max", torch.tensor(4.0))
        self.register_buffer("clip_factor_a_min", torch.tensor(4.0))

    def forward(self, x):
        import flatquant_b200.deploy as deploy
        if isinstance(x, deploy.PackedQuantizedTensor):
            return x
        if self.lac:
            cmax = float(torch.sigmoid(self.clip_factor_a_max.float()))
            cmin = float(torch.sigmoid(self.clip_factor_a_min.float()))
        else:
            cmax = cmin = float(self.input_clip_ratio)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        xq, s16, _ = ops.quantize_rows(x2, cmax, cmin, half_math=True)  # one pass: absmax, scale, round, pack
        return deploy.PackedQuantizedTensor(xq.view(*x.shape[:-1], -1), s16.view(-1, 1))
