"""4-bit storage helpers + float quantiser references (mirror of deploy/functional/quantization.py).

pack_i4 / unpack_i4 define the bit-exact storage contract: two's-complement nibbles, element 2j in
bits [3:0] and element 2j+1 in bits [7:4] of byte j.  CUDA tensors use the library kernels; CPU
tensors (offline checkpoint preparation, as in Linear4bit.from_float) use integer tensor ops."""
import torch

from ... import ops


def two_compl(x, bits: int):
    return torch.where(x < 0, x + 2 ** bits, x)


def get_minq_maxq(bits: int, sym: bool):
    if sym:
        hi = 2 ** (bits - 1) - 1
        return torch.tensor(-hi - 1), torch.tensor(hi)
    return torch.tensor(0), torch.tensor(2 ** bits - 1)


def sym_quant(x, scale, maxq):
    scale = scale.to(x.device)
    return torch.clamp(torch.round(x / scale), -(maxq + 1), maxq), scale


def sym_dequant(q, scale):
    return scale * q


def sym_quant_dequant(x, scale, maxq):
    return sym_dequant(*sym_quant(x, scale, maxq))


def asym_quant(x, scale, zero, maxq):
    scale, zero = scale.to(x.device), zero.to(x.device)
    return torch.clamp(torch.round(x / scale) + zero, 0, maxq), scale, zero


def asym_dequant(q, scale, zero):
    return scale * (q - zero)


def asym_quant_dequant(x, scale, zero, maxq):
    return asym_dequant(*asym_quant(x, scale, zero, maxq))


def pack_i4(q):
    """signed int tensor [rows, cols] in [-8, 7] -> uint8 [rows, cols/2]."""
    assert torch.is_signed(q), "The tensor to be packed should be signed int"
    assert bool(((q >= -8) & (q <= 7)).all()), "values outside the signed 4-bit range"
    q8 = q.to(torch.int8)
    if q8.is_cuda and q8.shape[-1] % 16 == 0:
        return ops.pack_i4(q8.contiguous())
    u = (q8.to(torch.int16) & 0xF).to(torch.uint8)
    return u[..., 0::2] | (u[..., 1::2] << 4)


def unpack_i4(x: torch.Tensor):
    """uint8 [..., c] -> int32 [..., 2c] (sign-extended)."""
    assert x.dtype == torch.uint8, "The tensor to be unpacked should be stored in uint8"
    if x.is_cuda and x.shape[-1] % 4 == 0:
        return ops.unpack_i4(x.contiguous())
    lo = (x & 0xF).to(torch.int32)
    hi = (x >> 4).to(torch.int32)
    both = torch.stack((lo - 16 * (lo >> 3), hi - 16 * (hi >> 3)), dim=-1)
    return both.reshape(*x.shape[:-1], x.shape[-1] * 2)
