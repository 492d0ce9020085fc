"""Quantiser modules with the reference's names, constructor arguments, attributes and buffers
(flatquant/quant_utils.py).  These are the fake-quant (float) definitions used for calibration,
weight preparation and as module-level API; the W4A4 inference arithmetic itself lives in the CUDA
library and is reached through FlatQuantizedLinear / deploy.nn."""
import torch


def round_ste(x):
    """round() with a straight-through gradient (flatquant/quant_utils.py:3-7)."""
    return x + (torch.round(x) - x).detach()


def get_qmin_qmax(bits, sym):
    """(q_max, q_min) of a `bits`-wide grid: symmetric [-2^(b-1), 2^(b-1)-1] as tensors, asymmetric [0, 2^b - 1]."""
    if not sym:
        return torch.tensor((1 << bits) - 1), 0
    top = (1 << (bits - 1)) - 1
    return torch.tensor(top), torch.tensor(-top - 1)


def _to_grid(x, scale, zero, lo, hi):
    """clamp(round_ste(x / scale) + zero, lo, hi): the one rounding step behind all four public helpers below."""
    q = round_ste(x / scale)
    return torch.clamp(q if zero is None else q + zero, lo, hi)


def sym_quant(x, scale, maxq):
    """-> (codes in [-(maxq+1), maxq], scale on x's device)  (quant_utils.py:19-23)"""
    s = scale.to(x.device)
    return _to_grid(x, s, None, -(maxq + 1), maxq), s


def asym_quant(x, scale, zero, maxq):
    """-> (codes in [0, maxq], scale, zero)  (quant_utils.py:33-38)"""
    s, z = scale.to(x.device), zero.to(x.device)
    return _to_grid(x, s, z, 0, maxq), s, z


def sym_dequant(q, scale):
    return q * scale


def asym_dequant(q, scale, zero):
    return (q - zero) * scale


def sym_quant_dequant(x, scale, maxq):
    return sym_dequant(*sym_quant(x, scale, maxq))


def asym_quant_dequant(x, scale, zero, maxq):
    return asym_dequant(*asym_quant(x, scale, zero, maxq))


class ActivationQuantizer(torch.nn.Module):
    """Per-token activation quantiser (flatquant/quant_utils.py:48-119): q_max/q_min, learnable
    clip factors `clip_factor_a_max/min` (raw, passed through a sigmoid) when lac=True."""

    def __init__(self, bits, sym=False, lac=False, groupsize=-1, clip_ratio=None):
        super().__init__()
        self.bits, self.sym, self.lac, self.groupsize = bits, sym, lac, groupsize
        self.q_max, self.q_min = get_qmin_qmax(bits, sym)
        self._clip_ratio = clip_ratio
        if lac:
            self.sigmoid = torch.nn.Sigmoid()
            self.clip_factor_a_max = torch.nn.Parameter(torch.full((1,), 4.0))
            self.clip_factor_a_min = torch.nn.Parameter(torch.full((1,), 4.0))
        self.enable = True

    def clip_factors(self):
        """Multiplicative (post-sigmoid) factors as Python floats, the form the C ABI takes."""
        if self.lac:
            return (float(torch.sigmoid(self.clip_factor_a_max.detach().float())),
                    float(torch.sigmoid(self.clip_factor_a_min.detach().float())))
        r = 1.0 if self._clip_ratio is None else float(self._clip_ratio)
        return r, r

    def forward(self, x, scale=None, zero=None):
        """`scale` / `zero`: precomputed quantisation parameters, and `groupsize` > 0: groups of that many channels
        quantised on their own -- the form of the vLLM copy (vllm_custom/.../fake_quant_utils.py:69-88, used for the
        grouped KV-cache quantisers); the main-repo class raises on groupsize > 0 (quant_utils.py:52-53)."""
        if self.bits == 16 or not self.enable:
            return x
        if self.groupsize > 0:
            shape = x.shape
            return self.fake_quant(x.reshape(-1, self.groupsize), scale, zero).reshape(shape)
        return self.fake_quant(x, scale, zero)

    def fake_quant(self, x, scale=None, zero=None):
        if scale is None or zero is None:
            scale, zero = self.get_scale_zero(x)
        qm = self.q_max.to(x)
        out = sym_quant_dequant(x, scale, qm) if self.sym else asym_quant_dequant(x, scale, zero, qm)
        return out.to(x.dtype)

    def get_scale_zero(self, x):
        qm = self.q_max.to(x)
        rows = x.reshape(-1, x.shape[-1])
        hi = rows.amax(1, keepdim=True).clamp(min=0)
        lo = rows.amin(1, keepdim=True).clamp(max=0)
        if self.lac:
            hi = hi * self.sigmoid(self.clip_factor_a_max)
            lo = lo * self.sigmoid(self.clip_factor_a_min)
        elif self._clip_ratio is not None:
            hi, lo = hi * self._clip_ratio, lo * self._clip_ratio
        if self.sym:
            m = torch.maximum(lo.abs(), hi)
            scale = m / qm
            scale = torch.where(m == 0, torch.ones_like(scale), scale)
            zero = torch.zeros_like(scale)
        else:
            dead = (lo == 0) & (hi == 0)
            lo = torch.where(dead, -torch.ones_like(lo), lo)
            hi = torch.where(dead, torch.ones_like(hi), hi)
            scale = (hi - lo) / qm
            zero = torch.round(-lo / scale)
        expand = lambda t: t.expand(-1, rows.shape[-1]).reshape(x.shape)  # noqa: E731
        return expand(scale), expand(zero)


class WeightQuantizer(torch.nn.Module):
    """GPTQ-style weight quantiser (flatquant/quant_utils.py:122-229): buffers maxq/scale/zero,
    configure(), find_params(), quantize(), enabled(), ready()."""

    def __init__(self, shape=1):
        super().__init__()
        self.register_buffer("maxq", torch.tensor(0))
        self.register_buffer("scale", torch.zeros(shape))
        self.register_buffer("zero", torch.zeros(shape))
        self.enable = True

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8):
        self.bits, self.perchannel, self.sym, self.mse = bits, perchannel, sym, mse
        self.norm, self.grid, self.maxshrink = norm, grid, maxshrink
        self.maxq = torch.tensor(2 ** (bits - 1) - 1 if sym else 2 ** bits - 1)

    def _range(self, rows):
        lo = rows.min(1)[0].clamp(max=0)
        hi = rows.max(1)[0].clamp(min=0)
        if self.sym:
            hi = torch.maximum(lo.abs(), hi).clamp(min=1e-5)
        else:
            dead = (lo == 0) & (hi == 0)
            lo = torch.where(dead, -torch.ones_like(lo), lo)
            hi = torch.where(dead, torch.ones_like(hi), hi)
        return lo, hi

    def _scale_zero(self, lo, hi, p=None):
        """p=None: the initial parameters (quant_utils.py:166-173, asym range clamped to 1e-5 BEFORE the division);
        p given: one step of the MSE grid (quant_utils.py:178-188, no clamp -- including its first step p == 1)."""
        if self.sym:
            s = (hi if p is None else p * hi) / self.maxq
            return s, torch.zeros_like(s)
        if p is None:
            s = (hi - lo).clamp(min=1e-5) / self.maxq
            return s, torch.round(-lo / s)
        s = (p * hi - p * lo) / self.maxq
        return s, torch.round(-p * lo / s)

    def find_params(self, x):
        if self.bits == 16 or not self.enable:
            return
        self.maxq = self.maxq.to(x.device)
        shape = x.shape
        rows = x.flatten(1) if self.perchannel else x.flatten().unsqueeze(0)
        lo, hi = self._range(rows)
        self.scale, self.zero = self._scale_zero(lo, hi)
        if self.mse:
            best = torch.full([rows.shape[0]], float("inf"), device=x.device)
            for i in range(int(self.maxshrink * self.grid)):
                s1, z1 = self._scale_zero(lo, hi, 1 - i / self.grid)
                if self.sym:
                    q = sym_quant_dequant(rows, s1.unsqueeze(1), self.maxq)
                else:
                    q = asym_quant_dequant(rows, s1.unsqueeze(1), z1.unsqueeze(1), self.maxq)
                err = (q - rows).abs_().pow_(self.norm).sum(1)
                better = err < best
                best = torch.where(better, err, best)
                self.scale = torch.where(better, s1, self.scale)
                self.zero = torch.where(better, z1, self.zero)
        if not self.perchannel:
            self.scale, self.zero = self.scale.repeat(shape[0]), self.zero.repeat(shape[0])
        bshape = [-1] + [1] * (len(shape) - 1)
        self.scale, self.zero = self.scale.reshape(bshape), self.zero.reshape(bshape)

    def quantize(self, x):
        if self.enable and self.ready() and self.bits < 16:
            if self.sym:
                return sym_quant_dequant(x, self.scale, self.maxq).to(x.dtype)
            return asym_quant_dequant(x, self.scale, self.zero, self.maxq).to(x.dtype)
        return x

    def forward(self, x):
        return self.quantize(x)

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)


def _set_state(model, kinds, enable):
    for m in model.modules():
        if isinstance(m, kinds):
            m.enable = enable
    return model


def set_quantizer_state(model, enable=True):
    return _set_state(model, (WeightQuantizer, ActivationQuantizer), enable)


def set_weight_quantizer_state(model, enable=True):
    return _set_state(model, WeightQuantizer, enable)


def set_act_quantizer_state(model, enable=True):
    return _set_state(model, ActivationQuantizer, enable)
