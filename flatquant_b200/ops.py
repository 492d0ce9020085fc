"""Tensor-level wrappers over the C ABI: argument checks, output/workspace allocation, stream.

Mirrors what the reference's pybind layer does (deploy/kernels/bindings.cpp:9-87: contiguity /
device checks, outputs allocated with torch::empty) -- torch is used for device memory and the
current stream only; all compute happens in libflatq_b200.so.
"""
from __future__ import annotations

import os

import torch

from . import _lib

F16, BF16 = 0, 1
_DT = {torch.float16: F16, torch.bfloat16: BF16}
_sm_checked = False
# opt-in until the split-KV decode kernel has run on hardware
_DECODE_SPLIT = os.environ.get("FLATQ_DECODE_SPLIT", "1") != "0"   # default since round 2 (hardware-verified); "0" opts out


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    global _sm_checked
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("flatquant_b200 ops are CUDA-only (sm_100a); got a CPU tensor. There is no CPU fallback.")
        if not t.is_contiguous():
            raise RuntimeError("flatquant_b200 ops need contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("flatquant_b200 ops need all tensors on the same GPU")
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        # like the reference's bindings (no device guard, bindings.cpp): kernels launch on the CURRENT device and stream
        raise RuntimeError(f"flatquant_b200 ops: tensors live on {dev} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; call under torch.cuda.device({dev.index})")
    if not _sm_checked:
        if not _lib.load().flatq_device_supported():
            raise RuntimeError("flatquant_b200 needs an sm_100 (B200) device: tcgen05/TMA kernels only")
        _sm_checked = True
    return dev


class _Workspace:
    """Grow-only per-device scratch so steady-state calls do not allocate (CUDA-graph friendly).  A buffer that is
    outgrown is retired, not freed: a captured CUDA graph may still hold its address, and handing that memory back to
    the allocator would let a later replay write into somebody else's tensor."""

    def __init__(self):
        self.buf = {}
        self.retired = []

    def get(self, dev, nbytes):
        b = self.buf.get(dev)
        if b is None or b.numel() < nbytes:
            if b is not None:
                self.retired.append(b)
            b = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
            self.buf[dev] = b
        return b


_ws = _Workspace()
_ws2 = _Workspace()
_ws3 = _Workspace()  # tile-prefix tables of grouped launches
_ws4 = _Workspace()  # partial softmax states of the split-KV decode attention


def pack_i4(q: torch.Tensor) -> torch.Tensor:
    """int8 codes [..., cols] in [-8,7] -> uint8 [..., cols/2] (deploy.functional.pack_i4)."""
    dev = _need_cuda(q)
    if q.dtype != torch.int8:
        q = q.to(torch.int8)
    cols = q.shape[-1]
    out = torch.empty(q.shape[:-1] + (cols // 2,), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().flatq_pack_i4(_ptr(q), _ptr(out), q.numel() // max(cols, 1), cols, _stream()), "pack_i4")
    return out


def unpack_i4(p: torch.Tensor) -> torch.Tensor:
    """uint8 [..., c] -> int32 [..., 2c] (deploy.functional.unpack_i4)."""
    dev = _need_cuda(p)
    assert p.dtype == torch.uint8, "The tensor to be unpacked should be stored in uint8"
    c = p.shape[-1]
    out = torch.empty(p.shape[:-1] + (2 * c,), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().flatq_unpack_i4(_ptr(p), _ptr(out), p.numel() // max(c, 1), c, _stream()), "unpack_i4")
    return out


def sym_quant(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """deploy._CUDA.sym_quant: x f16 [rows, cols], scale f16 [rows] -> packed [rows, cols/2]."""
    dev = _need_cuda(x, scale)
    rows, cols = x.shape
    if scale.numel() != rows:
        raise RuntimeError(f"sym_quant: scale has {scale.numel()} elements, expected {rows}")
    out = torch.empty((rows, cols // 2), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().flatq_sym_quant(_ptr(x), _ptr(scale), rows, cols, _ptr(out), _stream()), "sym_quant")
    return out


def sym_dequant(q: torch.Tensor, scale_row: torch.Tensor, scale_col: torch.Tensor, bias=None, out_dtype=torch.float16,
                ref_kernel_arith: bool = False) -> torch.Tensor:
    dev = _need_cuda(q, scale_row, scale_col, bias)
    rows, cols = q.shape
    if scale_row.numel() != rows or scale_col.numel() != cols:
        raise RuntimeError("sym_dequant: scale sizes do not match q")
    y = torch.empty((rows, cols), dtype=out_dtype, device=dev)
    _lib.check(_lib.load().flatq_sym_dequant(_ptr(q), _ptr(scale_row), _ptr(scale_col), _ptr(bias), rows, cols, _ptr(y),
                                             _DT[out_dtype], 1 if ref_kernel_arith else 0, _stream()), "sym_dequant")
    return y


def gemm_i4(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """deploy._CUDA.matmul: packed A [M, K/2] x packed B [N, K/2]^T -> int32 [M, N]."""
    dev = _need_cuda(a, b)
    m, kp = a.shape
    n = b.shape[0]
    if b.shape[1] != kp:
        raise RuntimeError("matmul: A and B disagree on K")
    lib = _lib.load()
    c = torch.empty((m, n), dtype=torch.int32, device=dev)
    if m == 0 or n == 0:
        return c
    nb = lib.flatq_gemm_i4_workspace_bytes(m, kp * 2)
    ws = _ws.get(dev, nb)
    _lib.check(lib.flatq_gemm_i4(_ptr(a), _ptr(b), m, n, kp * 2, _ptr(c), _ptr(ws), ws.numel(), _stream()), "gemm_i4")
    return c


def linear4bit(a: torch.Tensor, s_a: torch.Tensor, b: torch.Tensor, s_b: torch.Tensor, bias=None,
               out_dtype=torch.float16) -> torch.Tensor:
    """Linear4bit.forward in one launch chain: (A.B^T) * s_a[m] * s_b[n] (+bias)."""
    dev = _need_cuda(a, s_a, b, s_b, bias)
    m, kp = a.shape
    n = b.shape[0]
    lib = _lib.load()
    y = torch.empty((m, n), dtype=out_dtype, device=dev)
    ws = _ws.get(dev, lib.flatq_linear4bit_workspace_bytes(m, n, kp * 2))
    _lib.check(lib.flatq_linear4bit(_ptr(a), _ptr(s_a), _ptr(b), _ptr(s_b), _ptr(bias), m, n, kp * 2, _ptr(y),
                                    _DT[out_dtype], _ptr(ws), ws.numel(), _stream()), "linear4bit")
    return y


def kron_transform(x: torch.Tensor, left, right, n1: int, n2: int) -> torch.Tensor:
    """kronecker_matmul on fp16: x [..., n1*n2] -> same shape."""
    dev = _need_cuda(x, left, right)
    z = torch.empty_like(x)
    if left is None:  # rows are independent: group 16 of them per "token" so that no MMA row is padding
        rows = x.numel() // n2
        n1 = next(c for c in (16, 8, 4, 2, 1) if rows % c == 0)
    tokens = x.numel() // (n1 * n2)
    _lib.check(_lib.load().flatq_kron_transform(_ptr(x), _ptr(left), _ptr(right), tokens, n1, n2, _ptr(z), _stream()),
               "kron_transform")
    return z


def transform_quant(x: torch.Tensor, left, right, n1: int, n2: int, clip_max: float = 1.0, clip_min: float = 1.0):
    """OnlineTrans+Quantizer: returns (packed [tokens, K/2] u8, scales f16 [tokens], scales f32 [tokens])."""
    dev = _need_cuda(x, left, right)
    k = n1 * n2
    tokens = x.numel() // k
    xq = torch.empty((tokens, k // 2), dtype=torch.uint8, device=dev)
    s16 = torch.empty((tokens,), dtype=torch.float16, device=dev)
    s32 = torch.empty((tokens,), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().flatq_transform_quant(_ptr(x), _ptr(left), _ptr(right), tokens, n1, n2, float(clip_max),
                                                 float(clip_min), _ptr(xq), _ptr(s16), _ptr(s32), _stream()),
               "transform_quant")
    return xq, s16, s32


def quantize_rows(x: torch.Tensor, clip_max: float = 1.0, clip_min: float = 1.0, half_math: bool = True):
    dev = _need_cuda(x)
    k = x.shape[-1]
    tokens = x.numel() // k
    xq = torch.empty((tokens, k // 2), dtype=torch.uint8, device=dev)
    s16 = torch.empty((tokens,), dtype=torch.float16, device=dev)
    s32 = torch.empty((tokens,), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().flatq_quantize_rows(_ptr(x), tokens, k, float(clip_max), float(clip_min), int(half_math),
                                               _ptr(xq), _ptr(s16), _ptr(s32), _stream()), "quantize_rows")
    return xq, s16, s32


def linear_fused(x: torch.Tensor, left, right, n1: int, n2: int, clip_max: float, clip_min: float,
                 w_packed: torch.Tensor, s_w: torch.Tensor, bias=None, out_dtype=torch.float16) -> torch.Tensor:
    """The whole FlatLinear: transform -> per-token INT4 quant -> W4A4 GEMM -> dequant(+bias).

    x fp16 [..., K]; w_packed u8 [N, K/2]; s_w fp32 [N]; returns [..., N]."""
    dev = _need_cuda(x, left, right, w_packed, s_w, bias)
    k = n1 * n2
    if x.shape[-1] != k or w_packed.shape[1] * 2 != k:
        raise RuntimeError(f"linear_fused: K mismatch (x {x.shape[-1]}, factors {n1}x{n2}, W {w_packed.shape[1] * 2})")
    if x.dtype != torch.float16:
        raise RuntimeError("linear_fused: activations must be float16")
    if s_w.dtype != torch.float32:
        raise RuntimeError("linear_fused: s_w must be float32")
    m = x.numel() // k
    n = w_packed.shape[0]
    lib = _lib.load()
    y = torch.empty(x.shape[:-1] + (n,), dtype=out_dtype, device=dev)
    if m == 0:  # empty batch: nothing to launch (torch ops in the reference return an empty tensor too)
        return y
    ws = _ws.get(dev, lib.flatq_linear_fused_workspace_bytes(m, n, k))
    _lib.check(lib.flatq_linear_fused(_ptr(x), _ptr(left), _ptr(right), n1, n2, float(clip_max), float(clip_min),
                                      _ptr(w_packed), _ptr(s_w), _ptr(bias), m, n, _ptr(y), _DT[out_dtype], _ptr(ws),
                                      ws.numel(), _stream()), "linear_fused")
    return y


def linear_fused_silu(x: torch.Tensor, left, right, n1: int, n2: int, clip_max: float, clip_min: float,
                      w_packed_il: torch.Tensor, s_w_il: torch.Tensor) -> torch.Tensor:
    """act_fn(gate_proj(x)) * up_proj(x) from the fp16 activations: transform -> quantise -> GEMM over the row-interleaved
    [gate; up] weight -> SiLU * up (act_quant + gemm_w4a8_silu; one launch for one or two tokens).  Returns [..., N/2]."""
    dev = _need_cuda(x, left, right, w_packed_il, s_w_il)
    k = n1 * n2
    if x.shape[-1] != k or w_packed_il.shape[1] * 2 != k:
        raise RuntimeError(f"linear_fused_silu: K mismatch (x {x.shape[-1]}, factors {n1}x{n2}, W {w_packed_il.shape[1] * 2})")
    if x.dtype != torch.float16 or s_w_il.dtype != torch.float32:
        raise RuntimeError("linear_fused_silu: float16 activations, float32 weight scales")
    m = x.numel() // k
    n = w_packed_il.shape[0]
    y = torch.empty(x.shape[:-1] + (n // 2,), dtype=torch.float16, device=dev)
    if m == 0:
        return y
    lib = _lib.load()
    ws = _ws.get(dev, lib.flatq_linear_fused_silu_workspace_bytes(m, n, k))
    _lib.check(lib.flatq_linear_fused_silu(_ptr(x), _ptr(left), _ptr(right), n1, n2, float(clip_max), float(clip_min),
                                           _ptr(w_packed_il), _ptr(s_w_il), m, n, _ptr(y), _ptr(ws), ws.numel(), _stream()),
               "linear_fused_silu")
    return y


def act_quant(x: torch.Tensor, left, right, n1: int, n2: int, clip_max: float = 1.0, clip_min: float = 1.0, out=None):
    """Transform + per-token quantise into the GEMM operand format: (xq8 int8 [M, K] opaque, s_x f32 [M])."""
    dev = _need_cuda(x, left, right)
    k = n1 * n2
    m = x.numel() // k
    if out is None:
        out = (torch.empty((m, k), dtype=torch.int8, device=dev), torch.empty((m,), dtype=torch.float32, device=dev))
    xq8, sx = out
    _lib.check(_lib.load().flatq_act_quant(_ptr(x), _ptr(left), _ptr(right), n1, n2, float(clip_max), float(clip_min), m,
                                           _ptr(xq8), _ptr(sx), _stream()), "act_quant")
    return xq8, sx


def gemm_w4a8(xq8: torch.Tensor, s_x: torch.Tensor, w_packed: torch.Tensor, s_w: torch.Tensor, bias=None,
              out_dtype=torch.float16, out=None) -> torch.Tensor:
    """y[M, N] = (xq8 . W^T) * s_x[m] * s_w[n] (+bias) on the tcgen05 kernel."""
    dev = _need_cuda(xq8, s_x, w_packed, s_w, bias)
    m, k = xq8.shape
    n = w_packed.shape[0]
    y = out if out is not None else torch.empty((m, n), dtype=out_dtype, device=dev)
    lib = _lib.load()
    nb = lib.flatq_gemm_w4a8_workspace_bytes(m, n, k)
    ws = _ws2.get(dev, nb) if nb else None  # separate pool: xq8 / s_x usually live in the primary workspace
    _lib.check(lib.flatq_gemm_w4a8(_ptr(xq8), _ptr(s_x), _ptr(w_packed), _ptr(s_w), _ptr(bias), m, n, k, _ptr(y),
                                   _DT[out_dtype], _ptr(ws), ws.numel() if ws is not None else 0, _stream()), "gemm_w4a8")
    return y


def gemm_w4a8_silu(xq8: torch.Tensor, s_x: torch.Tensor, w_packed_il: torch.Tensor, s_w_il: torch.Tensor, out=None) -> torch.Tensor:
    """act_fn(gate_proj(x)) * up_proj(x) in one call: `w_packed_il` [2 * inter, K/2] / `s_w_il` hold gate and up rows
    interleaved (row 2j = gate j, row 2j + 1 = up j).  Returns [M, inter] fp16, bit-identical to gemm_w4a8 + silu_mul;
    for M > 128 the activation runs in the GEMM epilogue."""
    dev = _need_cuda(xq8, s_x, w_packed_il, s_w_il)
    m, k = xq8.shape
    n = w_packed_il.shape[0]
    y = out if out is not None else torch.empty((m, n // 2), dtype=torch.float16, device=dev)
    lib = _lib.load()
    nb = lib.flatq_gemm_w4a8_silu_workspace_bytes(m, n, k)
    ws = _ws2.get(dev, nb) if nb else None
    _lib.check(lib.flatq_gemm_w4a8_silu(_ptr(xq8), _ptr(s_x), _ptr(w_packed_il), _ptr(s_w_il), m, n, k, _ptr(y), _ptr(ws),
                                        ws.numel() if ws is not None else 0, _stream()), "gemm_w4a8_silu")
    return y


def gemm_w4a8_grouped(xq8: torch.Tensor, s_x: torch.Tensor, w_packed: torch.Tensor, s_w: torch.Tensor,
                      group_offsets: torch.Tensor, bias=None, out_dtype=torch.float16, out=None,
                      rows_per_group_hint: int = 0) -> torch.Tensor:
    """One launch for the FlatLinears of all routed experts of a stage (DeepSeek block, BASELINE configs[4]):
    rows [group_offsets[e], group_offsets[e + 1]) of xq8 [M, K] (tokens sorted by expert) are multiplied with
    w_packed[e] ([E, N, K/2]) and scaled by s_x[m] * s_w[e, n] (+ bias[e, n]) -> y [M, N].  `group_offsets` is an int32
    DEVICE tensor [E + 1]; nothing of the routing is read on the host.  Bit-identical to E gemm_w4a8 calls."""
    dev = _need_cuda(xq8, s_x, w_packed, s_w, group_offsets, bias)
    m, k = xq8.shape
    e, n, kp = w_packed.shape
    if kp * 2 != k or tuple(s_w.shape) != (e, n) or group_offsets.numel() != e + 1:
        raise RuntimeError(f"gemm_w4a8_grouped: shapes xq8 {tuple(xq8.shape)}, w_packed {tuple(w_packed.shape)}, "
                           f"s_w {tuple(s_w.shape)}, group_offsets {tuple(group_offsets.shape)}")
    if group_offsets.dtype != torch.int32 or s_x.dtype != torch.float32 or s_w.dtype != torch.float32:
        raise RuntimeError("gemm_w4a8_grouped: group_offsets must be int32, scales float32")
    y = out if out is not None else torch.empty((m, n), dtype=out_dtype, device=dev)
    if m == 0:
        return y
    lib = _lib.load()
    ws = _ws3.get(dev, lib.flatq_gemm_w4a8_grouped_workspace_bytes(e))
    _lib.check(lib.flatq_gemm_w4a8_grouped(_ptr(xq8), _ptr(s_x), _ptr(w_packed), _ptr(s_w), _ptr(bias),
                                           _ptr(group_offsets), e, m, n, k, int(rows_per_group_hint), _ptr(y),
                                           _DT[out_dtype], _ptr(ws), ws.numel(), _stream()), "gemm_w4a8_grouped")
    return y


# ---- elementwise glue of a decoder layer (default in the model stack; FLATQ_FUSED_GLUE=0 opts out, see csrc/glue.cuh)
FUSED_GLUE = os.environ.get("FLATQ_FUSED_GLUE", "1") != "0"   # default since round 2; "0" falls back to the torch expressions


def rmsnorm(x: torch.Tensor, weight=None, eps: float = 1e-5, mean_dim=None) -> torch.Tensor:
    """HF LlamaRMSNorm on fp16: weight * half(float(x) * rsqrt(sum(x^2) / mean_dim + eps)); weight None = weight-less."""
    dev = _need_cuda(x, weight)
    if x.dtype != torch.float16 or (weight is not None and weight.dtype != torch.float16):
        raise RuntimeError("rmsnorm: float16 activations and weight")
    k = x.shape[-1]
    y = torch.empty_like(x)
    _lib.check(_lib.load().flatq_rmsnorm(_ptr(x), _ptr(weight), x.numel() // k, k, float(mean_dim or k), float(eps), _ptr(y),
                                         _stream()), "rmsnorm")
    return y


def add_rmsnorm(h: torch.Tensor, y: torch.Tensor, weight=None, eps: float = 1e-5):
    """(h + y rounded to fp16, rmsnorm of that) in one pass: the residual add fused into the following norm."""
    dev = _need_cuda(h, y, weight)
    if h.dtype != torch.float16 or y.dtype != torch.float16 or (weight is not None and weight.dtype != torch.float16):
        raise RuntimeError("add_rmsnorm: float16 activations and weight")
    if h.shape != y.shape:
        raise RuntimeError("add_rmsnorm: shapes")
    k = h.shape[-1]
    h_out, normed = torch.empty_like(h), torch.empty_like(h)
    _lib.check(_lib.load().flatq_add_rmsnorm(_ptr(h), _ptr(y), _ptr(weight), h.numel() // k, k, float(k), float(eps), _ptr(h_out),
                                             _ptr(normed), _stream()), "add_rmsnorm")
    return h_out, normed


def rope_(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """In-place HF rotary embedding: x [..., tokens, heads, 128] contiguous fp16, cos / sin with tokens * 128 elements."""
    _need_cuda(x, cos, sin)
    heads, d = x.shape[-2], x.shape[-1]
    tokens = x.numel() // (heads * d)
    if x.dtype != torch.float16 or cos.dtype != torch.float16 or sin.dtype != torch.float16:
        raise RuntimeError("rope_: float16 tensors")
    if cos.numel() != tokens * d or sin.numel() != tokens * d:
        raise RuntimeError(f"rope_: cos / sin must hold tokens x head_dim = {tokens} x {d} values")
    _lib.check(_lib.load().flatq_rope(_ptr(x), _ptr(cos), _ptr(sin), tokens, heads, d, _stream()), "rope")
    return x


def silu_mul(gu: torch.Tensor) -> torch.Tensor:
    """gu [..., 2 * inter] = [gate | up] (fused gate/up GEMM output) -> silu(gate) * up [..., inter], fp16."""
    dev = _need_cuda(gu)
    if gu.dtype != torch.float16:
        raise RuntimeError("silu_mul: float16 activations")
    inter = gu.shape[-1] // 2
    out = torch.empty(gu.shape[:-1] + (inter,), dtype=torch.float16, device=dev)
    _lib.check(_lib.load().flatq_silu_mul(_ptr(gu), gu.numel() // (2 * inter), inter, _ptr(out), _stream()), "silu_mul")
    return out


def hadamard(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """out = scale * x . H_n over the last axis (n a power of two), fp16."""
    dev = _need_cuda(x)
    n = x.shape[-1]
    out = torch.empty_like(x)
    _lib.check(_lib.load().flatq_hadamard(_ptr(x), _ptr(out), x.numel() // n, n, float(scale), _stream()), "hadamard")
    return out


def hadamard_kn(x: torch.Tensor, hadK: torch.Tensor, K: int, scale: float) -> torch.Tensor:
    """matmul_hadU_cuda for n = K * 2^m: per row, x viewed as [K, n/K] -> hadK @ half(scale * FHT over the last axis)."""
    dev = _need_cuda(x, hadK)
    if x.dtype != torch.float16:
        raise RuntimeError("hadamard_kn: float16 activations")
    n = x.shape[-1]
    hk = hadK.to(device=dev, dtype=torch.float16).contiguous()
    if hk.shape != (K, K):
        raise RuntimeError(f"hadamard_kn: hadK must be [{K}, {K}]")
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().flatq_hadamard_kn(_ptr(x), _ptr(out), _ptr(hk), x.numel() // n, n, int(K), float(scale), _stream()),
               "hadamard_kn")
    return out


FUSED_KV = os.environ.get("FLATQ_FUSED_KV", "1") != "0"   # the model stack's RoPE -> transform -> KV quantise -> scatter in one launch


def _rows_view(t: torch.Tensor, heads: int):
    """t [tokens, heads, 128] or [b, s, heads, 128] whose (head, dim) block is dense and whose tokens are evenly strided
    (a column slice of the fused qkv GEMM output) -> (t, token stride in elements); anything else is copied."""
    d = t.shape[-1]
    if t.stride(-1) == 1 and t.stride(-2) == d and t.dim() in (3, 4):
        # the stride of a size-1 dimension is arbitrary ([b, 1, H, 128] decode views of a [b, N] GEMM output carry the
        # dense value there): take the token stride from the dimension that actually steps
        if t.dim() == 3 or t.shape[0] == 1:
            ld, even = t.stride(-3), True
        elif t.shape[1] == 1:
            ld, even = t.stride(0), True
        else:
            ld, even = t.stride(1), t.stride(0) == t.shape[1] * t.stride(1)
        if t.numel() // (heads * d) == 1:
            ld = max(ld, heads * d)
        if even and ld % 4 == 0 and ld >= heads * d:
            return t, ld
    return t.reshape(-1, heads, d).contiguous(), heads * d


def kv_fused(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot,
             layer: int, want_k: bool = False):
    """RoPE -> K / Q transform -> asymmetric 4-bit KV quantise -> page scatter in one launch (flatq_kv_fused).

    q [..., Hq, 128], k / v [..., Hkv, 128] fp16, possibly strided views of the fused qkv GEMM output (read in place);
    cos_t / sin_t fp16 [positions, 128] tables and pos int32 [tokens] (None: no RoPE); t_q / t_k fp16 [128, 128] or None.
    Returns (q_t [tokens, Hq, 128], k_t [tokens, Hkv, 128] or None)."""
    dev = _need_cuda(cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot)
    for t in (q, k, v):
        if not t.is_cuda or t.dtype != torch.float16:
            raise RuntimeError("kv_fused: q, k, v must be float16 CUDA tensors")
    hq, hkv, d = q.shape[-2], k.shape[-2], q.shape[-1]
    q, ld_q = _rows_view(q, hq)
    k, ld_k = _rows_view(k, hkv)
    v, ld_v = _rows_view(v, hkv)
    tokens = q.numel() // (hq * d)
    q_out = torch.empty((tokens, hq, d), dtype=torch.float16, device=dev)
    k_out = torch.empty((tokens, hkv, d), dtype=torch.float16, device=dev) if want_k else None
    n_layers, page_size = kv_data.shape[1], kv_data.shape[4]
    _lib.check(_lib.load().flatq_kv_fused(_ptr(q), _ptr(k), _ptr(v), ld_q, ld_k, ld_v, tokens, hq, hkv, d, _ptr(cos_t), _ptr(sin_t),
                                          _ptr(pos), _ptr(t_q), _ptr(t_k), _ptr(q_out), _ptr(k_out), _ptr(kv_data), _ptr(kv_param),
                                          _ptr(slot), int(layer), n_layers, page_size, _stream()), "kv_fused")
    return q_out, k_out


def kv_quant(x: torch.Tensor, kv_data=None, kv_param=None, slot=None, layer: int = 0, kv_select: int = 0):
    """Asymmetric 4-bit quantise + pack of K or V rows x [tokens, heads, 128] (fp16).

    Returns (packed u8 [tokens, heads, 64], param f16 [tokens, heads, 2] = (scale, zero)); with a paged cache
    (kv_data [pages, layers, 2, heads, page_size, 64], kv_param [..., page_size, 2], slot int32 [tokens]) the rows
    are also scattered into it."""
    dev = _need_cuda(x, kv_data, kv_param, slot)
    tokens, heads, d = x.shape
    q = torch.empty((tokens, heads, d // 2), dtype=torch.uint8, device=dev)
    par = torch.empty((tokens, heads, 2), dtype=torch.float16, device=dev)
    n_layers = kv_data.shape[1] if kv_data is not None else 1
    page_size = kv_data.shape[4] if kv_data is not None else 1
    _lib.check(_lib.load().flatq_kv_quant(_ptr(x), tokens, heads, d, _ptr(q), _ptr(par), _ptr(kv_data), _ptr(kv_param),
                                          _ptr(slot), layer, n_layers, kv_select, page_size, _stream()), "kv_quant")
    return q, par


FUSED_DECODE_FRONT = os.environ.get("FLATQ_DECODE_FRONT", "1") != "0"   # decode steps: kv_fused folded into the attention kernel


def kv_decode_fused_ok(q, k, cos_t, t_q, t_k) -> bool:
    """The shapes flatq_kv_decode_fused takes: one token per sequence, 2..16 query heads per kv head, head_dim 128, RoPE
    tables and both cache transforms present."""
    hq, hkv = q.shape[-2], k.shape[-2]
    return (FUSED_DECODE_FRONT and _DECODE_SPLIT and cos_t is not None and t_q is not None and t_k is not None and q.shape[-1] == 128
            and hq % hkv == 0 and 2 <= hq // hkv <= 16 and q.numel() // (hq * 128) == q.shape[0])


def kv_decode_fused(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot,
                    kv_indptr, kv_indices, last_page_len, layer: int, sm_scale=None, max_tokens_hint: int = 0) -> torch.Tensor:
    """A decode step's kv_fused + kv_decode_i4 in one kernel (+ the split merge): q [b, 1, Hq, 128] / k, v [b, 1, Hkv, 128]
    pre-RoPE views of the qkv GEMM output, read in place; the new token's K / V rows are quantised into slot[b] by the
    attention kernel itself.  Returns o [b, Hq, 128]."""
    dev = _need_cuda(cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot, kv_indptr, kv_indices, last_page_len)
    for t in (q, k, v):
        if not t.is_cuda or t.dtype != torch.float16:
            raise RuntimeError("kv_decode_fused: q, k, v must be float16 CUDA tensors")
    hq, hkv, d = q.shape[-2], k.shape[-2], q.shape[-1]
    batch = q.shape[0]
    q, ld_q = _rows_view(q, hq)
    k, ld_k = _rows_view(k, hkv)
    v, ld_v = _rows_view(v, hkv)
    o = torch.empty((batch, hq, d), dtype=torch.float16, device=dev)
    if sm_scale is None:
        sm_scale = d ** -0.5
    lib = _lib.load()
    ws = _ws4.get(dev, lib.flatq_kv_decode_i4_split_workspace_bytes(batch, hq))
    _lib.check(lib.flatq_kv_decode_fused(_ptr(q), _ptr(k), _ptr(v), ld_q, ld_k, ld_v, _ptr(cos_t), _ptr(sin_t), _ptr(pos), _ptr(t_q),
                                         _ptr(t_k), _ptr(slot), _ptr(o), _ptr(kv_data), _ptr(kv_param), _ptr(kv_indptr),
                                         _ptr(kv_indices), _ptr(last_page_len), batch, hq, hkv, d, kv_data.shape[4], layer,
                                         kv_data.shape[1], float(sm_scale), int(max_tokens_hint), _ptr(ws), ws.numel(), _stream()),
               "kv_decode_fused")
    return o


def kv_decode_i4(q: torch.Tensor, kv_data, kv_param, kv_indptr, kv_indices, last_page_len, layer: int = 0,
                 sm_scale=None, split=None, max_tokens_hint: int = 0) -> torch.Tensor:
    """Single-query attention over the paged INT4 cache: q [batch, q_heads, 128] fp16 -> o same shape.

    split=True (the default; FLATQ_DECODE_SPLIT=0 or split=False selects the single-pass kernel) takes the split-KV kernel
    (flatq_kv_decode_i4_split: sequence split over CTAs, one CTA per kv head serving its query-head group);
    `max_tokens_hint` is a host-side upper bound of the sequence lengths (0 = unknown)."""
    dev = _need_cuda(q, kv_data, kv_param, kv_indptr, kv_indices, last_page_len)
    batch, q_heads, d = q.shape
    o = torch.empty_like(q)
    if sm_scale is None:
        sm_scale = d ** -0.5
    lib = _lib.load()
    if split is None:
        split = _DECODE_SPLIT
    if split:
        ws = _ws4.get(dev, lib.flatq_kv_decode_i4_split_workspace_bytes(batch, q_heads))
        _lib.check(lib.flatq_kv_decode_i4_split(_ptr(q), _ptr(o), _ptr(kv_data), _ptr(kv_param), _ptr(kv_indptr),
                                                _ptr(kv_indices), _ptr(last_page_len), batch, q_heads, kv_data.shape[3], d,
                                                kv_data.shape[4], layer, kv_data.shape[1], float(sm_scale),
                                                int(max_tokens_hint), _ptr(ws), ws.numel(), _stream()), "kv_decode_i4_split")
        return o
    _lib.check(lib.flatq_kv_decode_i4(_ptr(q), _ptr(o), _ptr(kv_data), _ptr(kv_param), _ptr(kv_indptr),
                                      _ptr(kv_indices), _ptr(last_page_len), batch, q_heads, kv_data.shape[3], d,
                                      kv_data.shape[4], layer, kv_data.shape[1], float(sm_scale), _stream()),
               "kv_decode_i4")
    return o
