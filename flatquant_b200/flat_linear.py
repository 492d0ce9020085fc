"""FlatQuantizedLinear -- the operator the model wrappers call (mirror of flatquant/flat_linear.py).

Same constructor, attributes (`linear`, `weight_quantizer`, `act_quantizer`, `lwc`,
`clip_factor_w_max/min`, `_eval_mode`) and methods (`forward(hidden_states, qa_trans, out_trans)`,
`_ori_forward`, `_train_forward`, `_eval_forward`, `reparameterize`, `apply_trans`, `apply_wclip`).

Inference (`_eval_mode`): activations on a B200 go through ONE C-ABI call, `flatq_linear_fused`
(per-token INT4 quantise -> tcgen05 W4A4 GEMM -> dequant+bias); the 4-bit packed weight and its
per-channel scales are derived once from the (already reparameterised and RTN/GPTQ-rounded)
`linear.weight`.  `forward_fused(x, qa_trans)` additionally folds the caller's online Kronecker
transform into the same call.  Calibration (`_train_forward`) is torch autograd code, as in the
reference; it is not on the inference hot path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .flat_utils import kronecker_matmul
from .quant_utils import ActivationQuantizer, WeightQuantizer


class FlatQuantizedLinear(nn.Module):
    def __init__(self, args, linear: nn.Linear):
        super().__init__()
        self.args = args
        self.linear = linear
        self.weight_quantizer = WeightQuantizer()
        self.weight_quantizer.configure(args.w_bits, perchannel=True, sym=not args.w_asym, mse=False)
        self.act_quantizer = ActivationQuantizer(bits=args.a_bits, sym=not args.a_asym, lac=args.lac,
                                                 groupsize=args.a_groupsize)
        self.lwc = args.lwc
        if self.lwc:
            n_out = linear.weight.shape[0]
            self.clip_factor_w_max = nn.Parameter(torch.full((n_out, 1), 4.0))
            self.clip_factor_w_min = nn.Parameter(torch.full((n_out, 1), 4.0))
            self.sigmoid = nn.Sigmoid()
        self._eval_mode = False
        self._packed = None  # (key, w_packed u8 [N,K/2], s_w f32 [N])

    # ------------------------------------------------------------------ weight-side pieces (offline)
    def apply_wclip(self, weight):
        lo = weight.min(1, keepdim=True)[0] * self.sigmoid(self.clip_factor_w_min)
        hi = weight.max(1, keepdim=True)[0] * self.sigmoid(self.clip_factor_w_max)
        return torch.clamp(weight, min=lo, max=hi)

    def apply_trans(self, weight, qa_trans):
        if isinstance(qa_trans, list):  # [left_inv_t, right_inv_t] tensors (o_proj: llama_utils.py:279-281)
            return kronecker_matmul(weight, qa_trans[0].to(weight), qa_trans[1].to(weight))
        return qa_trans(weight, inv_t=True)

    def _transformed_weight(self, weight, qa_trans, out_trans):
        if qa_trans is not None:
            weight = self.apply_trans(weight, qa_trans)
        if self.lwc:
            weight = self.apply_wclip(weight)
        if out_trans is not None:
            weight = out_trans(weight.T).T
        return weight

    def reparameterize(self, qa_trans=None, out_trans=None):
        """Bake transform + clipping into linear.weight in fp64 and switch to eval mode
        (flat_linear.py:82-97)."""
        w = self.linear.weight.data
        new_w = self._transformed_weight(w.to(torch.float64), qa_trans, out_trans)  # fp64 -> torch route
        if out_trans is not None and self.linear.bias is not None:
            self.linear.bias.data = out_trans(self.linear.bias.data)
        self.linear.weight.data = new_w.detach().to(w.dtype)
        self._eval_mode = True
        self._packed = None

    # ------------------------------------------------------------------ forwards
    def _ori_forward(self, hidden_states):
        return self.linear(hidden_states)

    def _train_forward(self, hidden_states, qa_trans=None, out_trans=None):
        weight = self._transformed_weight(self.linear.weight.data, qa_trans, out_trans)
        self.weight_quantizer.find_params(weight)
        weight = self.weight_quantizer(weight)
        hidden_states = self.act_quantizer(hidden_states)
        bias = self.linear.bias
        if out_trans is not None and bias is not None:
            bias = out_trans(bias.data)
        return F.linear(hidden_states, weight, bias)

    def forward(self, hidden_states, qa_trans=None, out_trans=None):
        if self._eval_mode:
            return self._eval_forward(hidden_states)
        return self._train_forward(hidden_states, qa_trans=qa_trans, out_trans=out_trans)

    # ------------------------------------------------------------------ inference on the CUDA library
    def set_weight_scales(self, weight_scales):
        """Hand over the quantiser's true per-channel scales [N] / [N, 1] (what main.py holds in its `quantizers` dict
        after RTN / GPTQ, gptq_utils.py:272-310); pack_weights uses them instead of recovering the grid from the weight."""
        self._weight_scales = None if weight_scales is None else weight_scales.detach().reshape(-1).float().clone()
        self._packed = None

    def _grid_scales(self, wf):
        """Per-channel scale of an int-grid weight.  Order of trust: scales handed over by the caller, this module's own
        weight_quantizer when it holds per-channel parameters, then recovery from the weight itself: the extreme code of a
        row is +-7 after RTN but can be -8 or anything below 7 after GPTQ, so amax/7 is only the first candidate."""
        n = wf.shape[0]
        given = getattr(self, "_weight_scales", None)
        if given is not None and given.numel() == n:
            return given.to(wf).reshape(-1, 1)
        wq = self.weight_quantizer
        if getattr(wq, "scale", None) is not None and wq.scale.numel() == n and bool(wq.ready()):
            # only when it describes THIS weight (a calibration-time forward leaves the scale of the unrounded weight)
            s = wq.scale.detach().to(wf).reshape(-1, 1)
            codes = torch.round(wf / s)
            if torch.allclose(codes * s, wf, rtol=2e-3, atol=0) and codes.max() <= 7 and codes.min() >= -8:
                return s
        amax = wf.abs().amax(1, keepdim=True).clamp(min=1e-10)
        best = amax / 7
        todo = torch.ones(n, dtype=torch.bool, device=wf.device)
        for k in (7, 8, 6, 5, 4, 3, 2, 1):   # first candidate whose grid holds the whole row wins
            s = amax / k
            codes = torch.round(wf / s)
            ok = ((codes * s - wf).abs() <= 2e-3 * wf.abs()).all(1) & (codes.amax(1) <= 7) & (codes.amin(1) >= -8)
            take = todo & ok
            best = torch.where(take[:, None], s, best)
            todo &= ~ok
            if not bool(todo.any()):
                break
        return best

    def pack_weights(self, weight_scales=None):
        """4-bit pack the int-grid weight: returns (w_packed u8 [N, K/2], s_w f32 [N]) on the weight's device.

        `linear.weight` must already hold round(W/s)*s (RTN gptq_utils.py:296-304 or GPTQ); see _grid_scales for where
        the per-channel scales come from.  Codes outside [-8, 7] are rejected, never clamped."""
        w = self.linear.weight.data
        key = (w.data_ptr(), w._version, tuple(w.shape))
        if self._packed is not None and self._packed[0] == key and weight_scales is None:
            return self._packed[1], self._packed[2]
        wf = w.float()
        if weight_scales is not None:
            self.set_weight_scales(weight_scales)
        s = self._grid_scales(wf)
        codes = torch.round(wf / s)
        if not torch.allclose(codes * s, wf, rtol=2e-3, atol=0) or codes.max() > 7 or codes.min() < -8:
            raise RuntimeError("FlatQuantizedLinear: linear.weight is not on a per-channel INT4 grid; run RTN/GPTQ "
                               "weight quantisation (w_bits=4) before W4A4 inference")
        codes = codes.to(torch.int8)
        packed = ops.pack_i4(codes.contiguous()) if codes.is_cuda else _pack_cpu(codes)
        self._packed = (key, packed, s.reshape(-1).contiguous())
        return self._packed[1], self._packed[2]

    def _check_supported(self, x):
        aq = self.act_quantizer
        if aq.bits != 4 or not aq.sym or self.args.w_bits != 4 or self.args.w_asym:
            raise RuntimeError("flatquant_b200 implements the symmetric W4A4 path only")
        if not x.is_cuda:
            raise RuntimeError("FlatQuantizedLinear eval forward is CUDA-only (sm_100a kernels, no CPU fallback); "
                               "use oracle/ for CPU checking")

    def forward_fused(self, hidden_states, qa_trans=None):
        """Online transform + quantise + GEMM + dequant in one library call.  `qa_trans` is an eval-mode
        *DecomposeTransMatrix (diag already folded away), a [left, right] list, or None."""
        self._check_supported(hidden_states)
        left = right = None
        k = hidden_states.shape[-1]
        n1, n2 = 1, k
        if qa_trans is not None:
            if isinstance(qa_trans, (list, tuple)):
                left, right = qa_trans
            else:
                if getattr(qa_trans, "add_diag", False) and qa_trans.use_diag:
                    hidden_states = hidden_states * qa_trans.diag_scale.to(hidden_states)
                left, right = qa_trans.matrix_left, qa_trans.matrix_right
            n1, n2 = left.shape[0], right.shape[0]
            left = left.to(device=hidden_states.device, dtype=torch.float16).contiguous()
            right = right.to(device=hidden_states.device, dtype=torch.float16).contiguous()
        w_packed, s_w = self.pack_weights()
        cmax, cmin = self.act_quantizer.clip_factors()
        x = hidden_states
        out_dtype = x.dtype if x.dtype in (torch.float16, torch.bfloat16) else torch.float16
        x16 = x.to(torch.float16).contiguous()
        bias = self.linear.bias
        if bias is not None:
            bias = bias.data.to(out_dtype).contiguous()
        y = ops.linear_fused(x16, left, right, n1, n2, cmax, cmin, w_packed, s_w, bias, out_dtype)
        return y.to(x.dtype)

    def _eval_forward(self, hidden_states):
        if self.act_quantizer.bits == 16 or not self.act_quantizer.enable:
            return self.linear(hidden_states)
        return self.forward_fused(hidden_states, None)


def _pack_cpu(codes):
    """Offline format helper for CPU-resident checkpoints (not a compute fallback)."""
    u = (codes.to(torch.int16) & 0xF).to(torch.uint8)
    return (u[:, 0::2] | (u[:, 1::2] << 4)).contiguous()
