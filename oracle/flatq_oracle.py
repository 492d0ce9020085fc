"""CPU oracle for the FlatQuant W4A4 FlatLinear hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*: it may be imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``flatquant_b200/`` imports it; the product path is CUDA-only and
fails loudly without its extension.

It restates, with torch-CPU / numpy arithmetic, the algorithm of the reference's fake-quant
path and of its 4-bit storage format.  Every function cites the reference file:line it
follows (paths relative to the FlatQuant repository root).

Parity status: PINNED against the reference itself.  The reference ships no tests or golden
vectors of its own, so ``tests/golden/make_golden.py`` imports the
reference's Python modules in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# shape contract
# --------------------------------------------------------------------------------------


def get_decompose_dim(n: int) -> tuple[int, int]:
    """(n1, n2) with n1*n2 == n and n1 <= n2 as close as possible.

    Follows flatquant/function_utils.py:11-21 (copies: deploy/nn/online_trans.py:5-15,
    deploy/functional/online_trans.py:8-18): smallest a >= ceil(sqrt(n)) such that
    a*a - n is a perfect square b*b; result (a-b, a+b).
    """
    a = math.isqrt(n)
    if a * a < n:
        a += 1
    while True:
        d = a * a - n
        b = math.isqrt(d)
        if b * b == d:
            return a - b, a + b
        a += 1


# --------------------------------------------------------------------------------------
# online transform (a2/a3)
# --------------------------------------------------------------------------------------


def kronecker_matmul(x: torch.Tensor, left: torch.Tensor, right: torch.Tensor) -> torch.Tensor:
    """x @ kron(left, right) evaluated as left^T @ (X @ right) per token.

    flatquant/flat_utils.py:6-17.  X = x.reshape(-1, n1, n2) row-major.
    """
    n1, n2 = left.shape[0], right.shape[0]
    shp = x.shape
    xm = x.reshape(-1, n1, n2)
    y = torch.matmul(xm, right)
    z = torch.matmul(left.t(), y)
    return z.reshape(shp)


def single_matmul(x: torch.Tensor, mat: torch.Tensor) -> torch.Tensor:
    """[..., n] @ mat, the SVDSingle/InvSingle eval forward (flatquant/trans_utils.py:21-25)."""
    shp = x.shape
    return x.reshape(-1, mat.shape[0]).matmul(mat).reshape(shp)


def inverse_transpose(mat: torch.Tensor) -> torch.Tensor:
    """(mat^-1)^T in fp64, cast back (flatquant/function_utils.py:35-37, trans_utils.py:210-211)."""
    return mat.double().inverse().t().to(mat.dtype).contiguous()


# --------------------------------------------------------------------------------------
# activation quantizer (a4)
# --------------------------------------------------------------------------------------


def _sigmoid(v: float) -> float:
    return float(torch.sigmoid(torch.tensor(float(v), dtype=torch.float32)))


def act_scale(z: torch.Tensor, clip_max: float | None = None, clip_min: float | None = None,
              clip_ratio: float | None = None) -> torch.Tensor:
    """Per-token symmetric 4-bit scale, shape [tokens, 1], dtype of z.

    flatquant/quant_utils.py:77-107: hi=max(amax,0), lo=min(amin,0); with LAC both are
    multiplied by sigmoid(clip_factor) (clip_max/clip_min are the *raw* learnable factors);
    else by clip_ratio when given; s = max(|lo|, hi)/7, s=1 where that max is 0.
    """
    z2 = z.reshape(-1, z.shape[-1])
    hi = z2.amax(1, keepdim=True).clamp(min=0)
    lo = z2.amin(1, keepdim=True).clamp(max=0)
    if clip_max is not None:
        # the learnable factors are fp32 1-element parameters: with fp16/bf16 activations the
        # product (and hence scale, division and rounding) is promoted to fp32, as in the reference
        hi = hi * torch.sigmoid(torch.tensor([float(clip_max)], dtype=torch.float32))
        lo = lo * torch.sigmoid(torch.tensor([float(clip_min)], dtype=torch.float32))
    elif clip_ratio is not None:
        hi = hi * clip_ratio
        lo = lo * clip_ratio
    m = torch.maximum(lo.abs(), hi)
    s = m / 7
    s[m == 0] = 1
    return s


def sym_codes(z: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """clamp(round_half_even(z / s), -8, 7) as int8 (flatquant/quant_utils.py:19-22)."""
    z2 = z.reshape(-1, z.shape[-1])
    q = torch.clamp(torch.round(z2 / s), -8, 7)
    return q.to(torch.int8).reshape(z.shape)


def act_fake_quant(z: torch.Tensor, clip_max=None, clip_min=None, clip_ratio=None) -> torch.Tensor:
    """ActivationQuantizer.fake_quant, symmetric 4 bit (flatquant/quant_utils.py:71-82)."""
    s = act_scale(z, clip_max, clip_min, clip_ratio)
    z2 = z.reshape(-1, z.shape[-1])
    q = torch.clamp(torch.round(z2 / s), -8, 7)
    return (s * q).to(z.dtype).reshape(z.shape)


# --------------------------------------------------------------------------------------
# weight side, offline (a5/a7/a13)
# --------------------------------------------------------------------------------------


def weight_clip(w: torch.Tensor, cw_max: torch.Tensor, cw_min: torch.Tensor) -> torch.Tensor:
    """Learnable weight clipping: clamp rows to sigmoid-scaled row min/max (flat_linear.py:28-33)."""
    wmin = w.min(1, keepdim=True)[0] * torch.sigmoid(cw_min)
    wmax = w.max(1, keepdim=True)[0] * torch.sigmoid(cw_max)
    return torch.clamp(w, min=wmin, max=wmax)


def reparameterize_weight(w: torch.Tensor, left_inv_t: torch.Tensor | None, right_inv_t: torch.Tensor | None,
                          cw_max: torch.Tensor | None = None, cw_min: torch.Tensor | None = None,
                          out_mat: torch.Tensor | None = None) -> torch.Tensor:
    """FlatQuantizedLinear.reparameterize in fp64 (flatquant/flat_linear.py:82-97).

    Rows of W are transformed like activations with the inverse-transpose factors, then
    clipped, then optionally right-multiplied on the output side (v_proj out_trans).
    """
    dt = w.dtype
    w = w.double()
    if left_inv_t is not None:
        w = kronecker_matmul(w, left_inv_t.double(), right_inv_t.double())
    if cw_max is not None:
        w = weight_clip(w, cw_max.double(), cw_min.double())
    if out_mat is not None:
        w = single_matmul(w.t(), out_mat.double()).t()
    return w.to(dt)


def weight_rtn(w: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-output-channel symmetric RTN: (codes int8 [N,K], scale [N,1]).

    WeightQuantizer.find_params/quantize with perchannel, sym, no mse
    (flatquant/quant_utils.py:150-171,214-220) as driven by gptq_utils.py:296-304.
    """
    wmin = w.min(1)[0].clamp(max=0)
    wmax = w.max(1)[0].clamp(min=0)
    m = torch.maximum(wmin.abs(), wmax).clamp(min=1e-5)
    s = (m / 7).reshape(-1, 1)
    q = torch.clamp(torch.round(w / s), -8, 7).to(torch.int8)
    return q, s


# --------------------------------------------------------------------------------------
# 4-bit storage (a8) -- bit exact
# --------------------------------------------------------------------------------------


def pack_i4(q) -> np.ndarray:
    """[rows, cols] signed ints in [-8,7] -> uint8 [rows, cols/2].

    deploy/functional/quantization.py:4-5,49-56 and deploy/kernels/include/int4.h:86-98:
    two's-complement nibbles, element 2j in bits [3:0], element 2j+1 in bits [7:4].
    """
    q = np.asarray(q).astype(np.int16)
    assert q.min(initial=0) >= -8 and q.max(initial=0) <= 7
    u = (q & 0xF).astype(np.uint8)
    return (u[..., 0::2] | (u[..., 1::2] << 4)).astype(np.uint8)


def unpack_i4(p) -> np.ndarray:
    """Inverse of pack_i4 -> int32 (deploy/functional/quantization.py:59-82)."""
    p = np.asarray(p)
    assert p.dtype == np.uint8
    lo = (p & 0xF).astype(np.int32)
    hi = (p >> 4).astype(np.int32)
    lo[lo >= 8] -= 16
    hi[hi >= 8] -= 16
    out = np.empty(p.shape[:-1] + (p.shape[-1] * 2,), np.int32)
    out[..., 0::2] = lo
    out[..., 1::2] = hi
    return out


# --------------------------------------------------------------------------------------
# integer GEMM + dequant (a11/a12) and the fused linear (a6)
# --------------------------------------------------------------------------------------


def gemm_i4(xq, wq) -> np.ndarray:
    """int32 [M,N] = xq[M,K] . wq[N,K]^T (deploy/kernels/gemm.cu:7-46 semantics).

    Evaluated through fp64 BLAS: |acc| <= 64*K < 2^53, so the result is exact.
    """
    a = np.asarray(xq).astype(np.float64)
    b = np.asarray(wq).astype(np.float64)
    return np.rint(a @ b.T).astype(np.int32)


def sym_dequant(acc, s_row, s_col, bias=None) -> np.ndarray:
    """fp32 y = acc * s_row[m] * s_col[n] (+bias) -- the exact form of Linear4bit
    (deploy/nn/linear.py:48-54) == FlatQuantizedLinear._eval_forward (flat_linear.py:75-80)."""
    y = np.asarray(acc).astype(np.float32)
    y = y * np.asarray(s_row, np.float32).reshape(-1, 1) * np.asarray(s_col, np.float32).reshape(1, -1)
    if bias is not None:
        y = y + np.asarray(bias, np.float32).reshape(1, -1)
    return y


def sym_dequant_ref_kernel(acc, s_row, s_col) -> np.ndarray:
    """Bit-level model of the reference CUDA dequant (deploy/kernels/quant.cu:4-10,66-85):
    half(int(acc/10.0f) clamped to +-65176) then half multiplies, left to right:
    scale_row * scale_col * x * 10."""
    a = np.asarray(acc).astype(np.float32) / np.float32(10.0)
    t = np.clip(np.trunc(a), -65176, 65176).astype(np.float16)
    sr = np.asarray(s_row, np.float16).reshape(-1, 1)
    sc = np.asarray(s_col, np.float16).reshape(1, -1)
    y = (sr * sc).astype(np.float16)
    y = (y * t).astype(np.float16)
    return (y * np.float16(10.0)).astype(np.float16)


def sym_quant_half(x, scale) -> np.ndarray:
    """deploy.sym_quant codes (deploy/kernels/quant.cu:13-47): half division, round-to-nearest
    -even to int, clamp to [-8,7].  x fp16 [M,K], scale fp16 [M]."""
    x = np.asarray(x, np.float16)
    s = np.asarray(scale, np.float16).reshape(-1, 1)
    d = (x.astype(np.float32) / s.astype(np.float32)).astype(np.float16)
    return np.clip(np.rint(d.astype(np.float32)), -8, 7).astype(np.int8)


def flat_linear_int(x: torch.Tensor, left, right, clip_max, clip_min, wq, s_w, bias=None,
                    single: torch.Tensor | None = None):
    """Whole online path on integers: transform -> per-token scale -> codes -> int GEMM -> dequant.

    SURVEY.md appendix A; equals the fake-quant oracle `flat_linear_fake` to fp32 rounding.
    Returns dict with every intermediate so kernels can be checked stage by stage.
    """
    if single is not None:
        z = single_matmul(x, single)
    elif left is not None:
        z = kronecker_matmul(x, left.to(x.dtype), right.to(x.dtype))
    else:
        z = x
    z2 = z.reshape(-1, z.shape[-1])
    s_x = act_scale(z2, clip_max, clip_min)
    xq = sym_codes(z2, s_x)
    acc = gemm_i4(xq.numpy(), np.asarray(wq))
    y = sym_dequant(acc, s_x.float().numpy(), np.asarray(s_w, np.float32), bias)
    return {"z": z2, "s_x": s_x, "xq": xq.numpy(), "acc": acc, "y": y}


def flat_linear_fake(x: torch.Tensor, left, right, clip_max, clip_min, w_fq: torch.Tensor, bias=None):
    """The reference's fake-quant eval path: trans(x) -> fake-quant -> F.linear with the
    int-grid weight (trans_utils.py:97-101 + flat_linear.py:75-80)."""
    z = kronecker_matmul(x, left.to(x.dtype), right.to(x.dtype)) if left is not None else x
    zq = act_fake_quant(z, clip_max, clip_min)
    return torch.nn.functional.linear(zq, w_fq, bias)


# --------------------------------------------------------------------------------------
# qattention side path (a14)
# --------------------------------------------------------------------------------------


def hadamard_matrix(n: int) -> np.ndarray:
    """Sylvester H_n, n a power of two (the KAT of third-party/fast-hadamard-transform/tests/
    test_fast_hadamard_transform.py:31 uses scipy.linalg.hadamard, same construction)."""
    assert n & (n - 1) == 0 and n > 0
    h = np.ones((1, 1), np.float64)
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h


def hadamard_transform(x, scale: float = 1.0) -> np.ndarray:
    """out = scale * x @ H_n along the last axis, fp32 butterflies
    (third-party/fast-hadamard-transform/csrc/fast_hadamard_transform_cuda.cu:153-229 computes
    exactly this; H is symmetric so x@H == H@x per row)."""
    x = np.asarray(x, np.float32)
    n = x.shape[-1]
    y = x.reshape(-1, n).copy()
    h = 1
    while h < n:
        y = y.reshape(-1, n // (2 * h), 2, h)
        a = y[:, :, 0, :] + y[:, :, 1, :]
        b = y[:, :, 0, :] - y[:, :, 1, :]
        y = np.stack([a, b], axis=2).reshape(-1, n)
        h *= 2
    return (y * np.float32(scale)).reshape(x.shape)


def matmul_hadU_kn(x, had_k, scale: float, fp16_intermediate: bool = True) -> np.ndarray:
    """matmul_hadU_cuda for n = K * 2^m (deploy/functional/online_trans.py:144-151): x viewed as [..., K, n/K],
    hadamard_transform over the last axis with `scale` -- returned to torch as fp16, hence the rounding -- then
    `hadK @ input` over the K axis.  fp16_intermediate=False gives flatquant/hadamard_utils.py:89-110 (all fp32)."""
    x = np.asarray(x, np.float32)
    had_k = np.asarray(had_k, np.float32)
    k = had_k.shape[0]
    n = x.shape[-1]
    t = hadamard_transform(x.reshape(-1, k, n // k), scale)
    if fp16_intermediate:
        t = t.astype(np.float16).astype(np.float32)
    return np.einsum("ab,rbj->raj", had_k, t).reshape(x.shape)


def kv_asym_quant(x: torch.Tensor):
    """Per-(token, head) asymmetric 4-bit quantisation of K or V, non-LAC branch.

    deploy/transformers/kv_cache.py:11-51 (`lac=False`): s=(max-min).clamp(1e-5)/15,
    zero=-min, q=clamp(round((x+zero)/s),0,15); packed element 2j -> low nibble.
    Returns (packed uint8 [..., d/2], scale [...,1], zero [...,1]) in x's dtype.
    """
    xmax = x.amax(-1, keepdim=True)
    xmin = x.amin(-1, keepdim=True)
    s = (xmax - xmin).clamp(min=1e-5) / 15
    zero = -xmin
    q = torch.clamp(torch.round((x + zero) / s), 0, 15).to(torch.uint8)
    packed = q[..., 0::2] | (q[..., 1::2] << 4)
    return packed, s, zero


def kv_asym_dequant(packed: torch.Tensor, s: torch.Tensor, zero: torch.Tensor) -> torch.Tensor:
    """x_hat = nibble * s - zero (deploy/transformers/kv_cache.py:54-61, non-LAC;
    deploy/kernels/include/flashinfer/quantization.cuh:59-84)."""
    q = torch.stack((packed & 0xF, (packed >> 4) & 0xF), dim=-1).reshape(*packed.shape[:-1], -1)
    return q.to(s.dtype) * s - zero


def decode_attention_i4(q: torch.Tensor, k_packed, k_s, k_z, v_packed, v_s, v_z) -> torch.Tensor:
    """Single-query attention over a dequantised INT4 KV cache, fp32.

    q [B,H,D]; k/v packed [B,T,H,D/2] with params [B,T,H,1]
    (deploy/kernels/include/flashinfer/decode.cuh:480-689: softmax(q.K^T/sqrt(D)).V).
    """
    k = kv_asym_dequant(k_packed, k_s, k_z).float()
    v = kv_asym_dequant(v_packed, v_s, v_z).float()
    d = q.shape[-1]
    logits = torch.einsum("bhd,bthd->bht", q.float(), k) / math.sqrt(d)
    p = torch.softmax(logits, dim=-1)
    return torch.einsum("bht,bthd->bhd", p, v)
