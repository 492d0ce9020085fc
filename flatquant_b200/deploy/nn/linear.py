"""Linear4bit (mirror of deploy/nn/linear.py:22-77): buffers `weight` u8 [N, K/2], `weight_scales`
[N, 1], optional `bias`; forward takes a PackedQuantizedTensor and returns fp16."""
import torch

from ... import ops


class Linear4bit(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=False, dtype=torch.float16):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("weight_scales", torch.zeros((out_features, 1), requires_grad=False))
        self.register_buffer("weight", torch.randint(1, 7, (out_features, in_features // 2), dtype=torch.uint8,
                                                     requires_grad=False))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features,), dtype=dtype))
        else:
            self.bias = None

    def forward(self, x):
        """GEMM + dequant (+bias) in one library call (reference: deploy.matmul -> deploy.sym_dequant
        -> `+ bias`, linear.py:48-54; same fp32 arithmetic, one rounding)."""
        import flatquant_b200.deploy as deploy
        assert type(x) == deploy.PackedQuantizedTensor  # quantized input is given
        q, s = x.quantized_x, x.scales_x
        lead = q.shape[:-1]
        q2 = q.reshape(-1, q.shape[-1]).contiguous()
        out_dtype = self.weight_scales.dtype if self.weight_scales.dtype in (torch.float16, torch.bfloat16) else torch.float16
        y = ops.linear4bit(q2, s.reshape(-1).to(torch.float16).contiguous(), self.weight.contiguous(),
                           self.weight_scales.reshape(-1).to(torch.float16).contiguous(),
                           None if self.bias is None else self.bias.to(out_dtype).contiguous(), out_dtype)
        return y.view(*lead, -1)

    @staticmethod
    def from_float(module: torch.nn.Linear, weight_scales=None):
        """Build from an fp16 Linear whose weight is s*round(W/s): packs round(W/weight_scales)
        (linear.py:56-77)."""
        from ..functional.quantization import pack_i4
        w = module.weight.data
        m = Linear4bit(module.in_features, module.out_features, bias=module.bias is not None, dtype=w.dtype).to(w.dtype)
        if weight_scales is not None:
            assert weight_scales.shape == (module.out_features, 1), "weight_scales should have shape (out_features, 1)"
            m.weight_scales.copy_(weight_scales.to(w.dtype))
            codes = (w.to(weight_scales.device) / weight_scales).round().to(torch.int8)
            m.weight.copy_(pack_i4(codes).cpu())
            if module.bias is not None:
                m.bias.copy_(module.bias)
        return m
