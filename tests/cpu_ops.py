"""TEST INFRASTRUCTURE: CPU stand-ins for `flatquant_b200.ops`, built on the oracle (oracle/flatq_oracle.py).

Host-side code above the C ABI (model wrappers, loaders, KV-cache bookkeeping, tensor-parallel glue) is plain Python
and can be checked without a GPU by swapping the CUDA entry points for these functions (`install(monkeypatch)`).
They take and return CPU tensors with the shapes / dtypes of the real ops.  Nothing under flatquant_b200/ imports
this file; on the GPU box the `-m gpu` tests run the same host code on the real kernels and compare both.

`emulate_fp16=True` places fp16 roundings where the kernels place them (factors and activations fp16, intermediate and
transformed value rounded to fp16, tests/test_gpu_parity.py::_emulated_codes); False is the pure fp32 oracle."""
import torch

from oracle import flatq_oracle as O

EMULATE_FP16 = True

# ---------------------------------------------------------------------------------------------------------------------
# STRICT: every stand-in first applies the argument checks the real call would apply -- the ones of flatquant_b200/ops.py
# (contiguity) and of the C ABI (flatquant_b200/csrc/api.cu: shapes, divisibility, 16-byte alignment, supported factor
# sizes) plus the element types the kernels read through raw pointers (an fp16 tensor handed to a `const float*` is
# silent garbage on the GPU and invisible to an oracle-backed stand-in that calls .float()).  A host-logic test that
# passes here therefore only makes calls the CUDA library accepts.  Activation dtype checks apply in the deployed mode
# (EMULATE_FP16); the fp32 wiring tests run the same code with ACT_DTYPE = float32 on purpose.
STRICT = True
CALLS = []  # (op, config) of every call, for tests that want to see which kernel configurations a model exercises


class AbiViolation(AssertionError):
    pass


def _req(cond, msg):
    if STRICT and not cond:
        raise AbiViolation(msg)


def _contig(**ts):
    for name, t in ts.items():
        if t is not None:
            _req(t.is_contiguous(), f"{name}: ops._need_cuda needs contiguous tensors")


def _al16(t, elem_bytes, name):
    """16-byte alignment of the address the GPU would see (allocations are >= 256-byte aligned; the offset counts)."""
    if t is not None:
        _req((t.storage_offset() * elem_bytes) % 16 == 0, f"{name}: not 16-byte aligned (offset {t.storage_offset()} elements)")


def _act_dtype(t, name):
    if EMULATE_FP16 and t is not None:
        _req(t.dtype == torch.float16, f"{name}: the kernels read fp16, got {t.dtype}")


def _dtype(t, want, name):
    if t is not None:
        _req(t.dtype == want, f"{name}: the kernel reads {want}, got {t.dtype}")


def _check_kron(tokens, n1, n2, left, right, quant, z_out, x, out8):
    """launch_kron / launch_kron_tc of api.cu: which kernel a (n1, n2) request lands on, or a refusal."""
    _req(tokens > 0, "kron: no tokens (api.cu launch_kron)")
    _al16(x, 2, "kron x")
    _act_dtype(x, "kron x"), _act_dtype(left, "left factor"), _act_dtype(right, "right factor")
    if left is not None:
        _req(tuple(left.shape) == (n1, n1), f"left factor {tuple(left.shape)} != ({n1}, {n1})")
    if right is not None:
        _req(tuple(right.shape) == (n2, n2), f"right factor {tuple(right.shape)} != ({n2}, {n2})")
    if quant and not z_out and (n1 * n2) % 32 == 0:
        if left is not None and right is None and n2 == 128 and n1 in (16, 32, 64):
            return "tcgen05"   # launch_kron_ts: left-only over the heads axis runs with R = I (exact)
        if (n2 == 64 and (32 <= n1 <= 64)) or (n2 == 128 and 32 < n1 <= 128):
            return "tcgen05"
    _req(1 <= n1 <= 256 and 16 <= n2 <= 256 and n2 % 16 == 0,
         f"unsupported Kronecker factor sizes ({n1}, {n2}): need n2 % 16 == 0 and n1, n2 <= 256")
    p1 = (n1 + 15) // 16 * 16
    wpt, kt2 = p1 // 16, n2 // 16
    _req(wpt * 32 <= (512 if kt2 <= 8 else 256), f"Kronecker factors ({n1}, {n2}) need {wpt} warps per token")
    smem = ((p1 * (p1 + 8) * 2 if left is not None else 0) + (n2 * (n2 + 8) * 2 if right is not None else 0) +
            p1 * (n2 + 8) * 2 + 16 * 2 * 4 + 16)  # one token per CTA
    _req(smem <= 227 * 1024, f"Kronecker factors ({n1}, {n2}) exceed shared memory")
    return "hmma"


def _check_rowquant(x, tokens, k):
    _req(tokens > 0 and k > 0 and k % 8 == 0, f"quantize_rows: K={k} must be a multiple of 8 and tokens > 0")
    _al16(x, 2, "rowquant x"), _act_dtype(x, "rowquant x")


def _check_gemm(m, n, k, xq8, s_x, w_packed, s_w, bias, out, out_dtype):
    _req(m > 0 and n > 0 and k > 0, f"empty GEMM {m}x{n}x{k} (api.cu launch_gemm)")
    _req(k % 32 == 0, f"gemm_w4a8: K={k} must be a multiple of 32")
    _req(k <= 65536, "GEMM too large")
    _req(n % 2 == 0, f"N={n} must be even for 16-bit outputs")
    _req((m + 256) * n < (1 << 32), "output exceeds the 32-bit element index of the store path")
    _req(tuple(w_packed.shape) == (n, k // 2), f"w_packed {tuple(w_packed.shape)} != ({n}, {k // 2})")
    _dtype(xq8, torch.int8, "xq8"), _dtype(w_packed, torch.uint8, "w_packed")
    _dtype(s_x, torch.float32, "s_x"), _dtype(s_w, torch.float32, "s_w")
    _req(s_x.numel() == m and s_w.numel() == n, "scale sizes")
    _al16(xq8, 1, "xq8 (TMA)"), _al16(w_packed, 1, "w_packed (TMA)")
    _req(out_dtype in (torch.float16, torch.bfloat16) or not EMULATE_FP16, f"gemm_w4a8: out_dtype {out_dtype}")
    if bias is not None:
        _req(bias.numel() == n, "bias size")
        if EMULATE_FP16:
            _dtype(bias, out_dtype, "bias (read as the output type)")
    if out is not None:
        _req(tuple(out.shape) == (m, n) and out.is_contiguous(), "out must be a contiguous [M, N] tensor")
        _req((out.storage_offset() * 2) % 4 == 0, "out: 32-bit stores")


def _h(t):
    return t.half().float() if EMULATE_FP16 else t.float()


# the factor pairs served by the second tcgen05 design (csrc/kron_quant_ts.cuh, launch_kron_ts of api.cu)
TS_SHAPES = {(64, 64), (56, 64), (32, 64), (64, 128), (112, 128), (128, 128)}


def _transform(x2, left, right, n1, n2, right_first=False):
    """The kernels' rounding points: one fp16 rounding after each factor.  Left factor first -- (L^T X) R, the order of
    the deployed reference (kron_matmul.py:62-76), of kron_quant_ts.cuh and of the legacy-MMA kernel kron_quant.cuh --
    except on the first tcgen05 design (kron_quant_tc.cuh: L^T (X R)), which still serves a few padded shapes."""
    xm = _h(x2).reshape(-1, n1, n2)
    steps = [("r", right), ("l", left)] if right_first else [("l", left), ("r", right)]
    for kind, m in steps:
        if m is None:
            continue
        xm = _h(torch.matmul(xm, _h(m))) if kind == "r" else _h(torch.matmul(_h(m).t(), xm))
    return xm.reshape(x2.shape[0], n1 * n2)


def kron_transform(x, left, right, n1, n2):
    _contig(x=x, left=left, right=right)
    if left is None:
        rows = x.numel() // n2
        g = next(c for c in (16, 8, 4, 2, 1) if rows % c == 0)  # the grouping ops.kron_transform applies
        CALLS.append(("kron_transform", (g, n2, False, right is not None)))
        _check_kron(rows // g, g, n2, None, right, False, True, x, False)
        n1 = 1
    else:
        CALLS.append(("kron_transform", (n1, n2, True, right is not None)))
        _check_kron(x.numel() // (n1 * n2), n1, n2, left, right, False, True, x, False)
    return _transform(x.reshape(-1, n1 * n2), left, right, n1, n2).to(x.dtype).reshape(x.shape)


def _scale_codes(z, clip_max, clip_min):
    hi = z.amax(1).clamp(min=0) * torch.tensor(float(clip_max))
    lo = z.amin(1).clamp(max=0) * torch.tensor(float(clip_min))
    m = torch.maximum(lo.abs(), hi)
    s = torch.where(m == 0, torch.ones_like(m), m / 7)
    return torch.clamp(torch.round(z / s[:, None]), -8, 7).to(torch.int8), s.float()


def act_quant(x, left, right, n1, n2, clip_max=1.0, clip_min=1.0, out=None):
    """(codes int8 [M, K] -- natural K order, the format is opaque to callers --, s_x fp32 [M])"""
    k = n1 * n2
    _contig(x=x, left=left, right=right)
    _req(k % 32 == 0, f"act_quant: K={k} must be a multiple of 32")
    m = x.numel() // k
    kernel = None
    if left is None and right is None:
        _check_rowquant(x, m, k)
        CALLS.append(("act_quant", ("rowquant", k)))
    else:
        kernel = _check_kron(m, n1, n2, left, right, True, False, x, True)
        CALLS.append(("act_quant", (kernel, n1, n2, left is not None, right is not None)))
    if out is not None:
        _contig(xq8=out[0], s_x=out[1])
        _dtype(out[0], torch.int8, "act_quant out codes"), _dtype(out[1], torch.float32, "act_quant out scales")
        _al16(out[0], 1, "act_quant xq8")
    first_design = (left is not None or right is not None) and kernel == "tcgen05" and not (
        (n1, n2) in TS_SHAPES and left is not None and right is not None)
    z = _transform(x.reshape(-1, k), left, right, n1, n2, right_first=first_design)
    q, s = _scale_codes(z, clip_max, clip_min)
    if out is not None:
        out[0].copy_(q)
        out[1].copy_(s)
        return out
    return q, s


def transform_quant(x, left, right, n1, n2, clip_max=1.0, clip_min=1.0):
    _req((n1 * n2) % 32 == 0 or (left is None and right is None), "transform_quant: packed rows must be 16-byte multiples")
    q, s = act_quant(x, left, right, n1, n2, clip_max, clip_min)
    return torch.from_numpy(O.pack_i4(q.numpy())), s.half(), s


def quantize_rows(x, clip_max=1.0, clip_min=1.0, half_math=True):
    _contig(x=x)
    _check_rowquant(x, x.numel() // x.shape[-1], x.shape[-1])
    q, s = _scale_codes(x.reshape(-1, x.shape[-1]).float(), clip_max, clip_min)
    return torch.from_numpy(O.pack_i4(q.numpy())), s.half(), s


def gemm_w4a8(xq8, s_x, w_packed, s_w, bias=None, out_dtype=torch.float16, out=None):
    _contig(xq8=xq8, s_x=s_x, w_packed=w_packed, s_w=s_w, bias=bias)
    m, k = xq8.shape
    n = w_packed.shape[0]
    _check_gemm(m, n, k, xq8, s_x, w_packed, s_w, bias, out, out_dtype)
    CALLS.append(("gemm_w4a8", ("gemv" if m <= 2 and m * k <= 96 * 1024 else "skinny" if m <= 32 and k % 128 == 0 else "split-k" if m <= 64 else
                                "1sm-128" if m <= 128 else "2sm", m, n, k)))
    wq = O.unpack_i4(w_packed.numpy())
    acc = O.gemm_i4(xq8.numpy(), wq)
    y = O.sym_dequant(acc, s_x.numpy(), s_w.float().numpy(), None if bias is None else bias.float().numpy())
    y = torch.from_numpy(y).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_w4a8_silu(xq8, s_x, w_packed_il, s_w_il, out=None):
    n = w_packed_il.shape[0]
    _req(n % 16 == 0, f"gemm_w4a8_silu: N={n} must be a multiple of 16")
    gu = gemm_w4a8(xq8, s_x, w_packed_il, s_w_il)
    CALLS[-1] = ("gemm_w4a8_silu", CALLS[-1][1])
    y = torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2]
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_w4a8_grouped(xq8, s_x, w_packed, s_w, group_offsets, bias=None, out_dtype=torch.float16, out=None,
                      rows_per_group_hint=0):
    _contig(xq8=xq8, s_x=s_x, w_packed=w_packed, s_w=s_w, group_offsets=group_offsets, bias=bias)
    m, k = xq8.shape
    e, n, kp = w_packed.shape
    _req(kp * 2 == k and tuple(s_w.shape) == (e, n) and group_offsets.numel() == e + 1, "gemm_w4a8_grouped: shapes")
    _dtype(group_offsets, torch.int32, "group_offsets"), _dtype(xq8, torch.int8, "xq8"), _dtype(w_packed, torch.uint8, "w_packed")
    _dtype(s_x, torch.float32, "s_x"), _dtype(s_w, torch.float32, "s_w")
    _req(m > 0 and n > 0 and k % 32 == 0 and k <= 65536 and n % 2 == 0, f"gemm_w4a8_grouped: {m}x{n}x{k}")
    _req((m + 256) * n < (1 << 32) and e * n < (1 << 31), "gemm_w4a8_grouped: 32-bit index of the store path")
    _al16(xq8, 1, "xq8 (TMA)"), _al16(w_packed, 1, "w_packed (TMA)")
    off = group_offsets.tolist()
    _req(off[0] == 0 and off[-1] <= m and all(a <= b for a, b in zip(off, off[1:])), f"group_offsets {off[:4]}...")
    CALLS.append(("gemm_w4a8_grouped", (e, m, n, k)))
    y = out if out is not None else torch.zeros((m, n), dtype=out_dtype)
    for g in range(e):
        if off[g + 1] > off[g]:
            r = slice(off[g], off[g + 1])
            y[r] = gemm_w4a8(xq8[r], s_x[r], w_packed[g], s_w[g], None if bias is None else bias[g], out_dtype)
            CALLS.pop()
    return y


def linear_fused(x, left, right, n1, n2, clip_max, clip_min, w_packed, s_w, bias=None, out_dtype=torch.float16):
    k = n1 * n2
    _contig(x=x, left=left, right=right, w_packed=w_packed, s_w=s_w, bias=bias)
    _req(x.shape[-1] == k and w_packed.shape[1] * 2 == k, "linear_fused: K mismatch")
    _act_dtype(x, "linear_fused x"), _dtype(s_w, torch.float32, "linear_fused s_w")
    if x.numel() == 0:
        return torch.empty(x.shape[:-1] + (w_packed.shape[0],), dtype=out_dtype)
    q, s = act_quant(x.reshape(-1, k), left, right, n1, n2, clip_max, clip_min)
    y = gemm_w4a8(q, s, w_packed, s_w, bias, out_dtype).reshape(x.shape[:-1] + (w_packed.shape[0],))
    _tag_decode(x.numel() // k, w_packed.shape[0], n1, n2, left, right, False)
    return y


def _decode_fused(m, n1, n2, left, right, silu):
    """mirrors flat_decode_ok of csrc/api.cu: the shapes that run as ONE launch (csrc/flatlinear_decode.cuh)"""
    k = n1 * n2
    if not (1 <= m <= 2) or (left is None and right is None) or n1 % 16 or not (16 <= n1 <= 128) or n2 not in (64, 128) or k % 1024:
        return False
    cpl = k // 1024
    return cpl in (2, 4) if n2 == 64 else (cpl == 8 if silu else cpl in (4, 8))


def _tag_decode(m, n, n1, n2, left, right, silu):
    if _decode_fused(m, n1, n2, left, right, silu):   # replace the (act_quant, gemm) pair just recorded
        del CALLS[-2:]
        CALLS.append(("linear_fused_decode", (m, n, n1 * n2, n1, n2, silu)))


def linear_fused_silu(x, left, right, n1, n2, clip_max, clip_min, w_packed_il, s_w_il):
    k = n1 * n2
    _contig(x=x, left=left, right=right, w_packed_il=w_packed_il, s_w_il=s_w_il)
    _req(x.shape[-1] == k and w_packed_il.shape[1] * 2 == k, "linear_fused_silu: K mismatch")
    _act_dtype(x, "linear_fused_silu x"), _dtype(s_w_il, torch.float32, "linear_fused_silu s_w")
    n = w_packed_il.shape[0]
    if x.numel() == 0:
        return torch.empty(x.shape[:-1] + (n // 2,), dtype=x.dtype)
    q, s = act_quant(x.reshape(-1, k), left, right, n1, n2, clip_max, clip_min)
    y = gemm_w4a8_silu(q, s, w_packed_il, s_w_il).reshape(x.shape[:-1] + (n // 2,))
    _tag_decode(x.numel() // k, n, n1, n2, left, right, True)
    return y


def pack_i4(q):
    _contig(q=q)
    _req(q.shape[-1] % 16 == 0, f"pack_i4: cols={q.shape[-1]} must be a multiple of 16")
    _al16(q, 1, "pack_i4 q")
    return torch.from_numpy(O.pack_i4(q.numpy()))


def rmsnorm(x, weight=None, eps=1e-5, mean_dim=None):
    """the torch expression flatq_rmsnorm replaces (HF LlamaRMSNorm / deploy.nn.RMSNorm)"""
    _contig(x=x, weight=weight), _al16(x, 2, "rmsnorm x"), _act_dtype(x, "rmsnorm x"), _act_dtype(weight, "rmsnorm weight")
    _req(x.shape[-1] % 8 == 0, "rmsnorm: K must be a multiple of 8")
    xf = x.float()
    n = (xf * torch.rsqrt(xf.pow(2).sum(-1, keepdim=True) / float(mean_dim or x.shape[-1]) + eps)).to(x.dtype)
    return n if weight is None else weight * n


def add_rmsnorm(h, y, weight=None, eps=1e-5):
    """torch's `h = h + y` (rounded to the activation dtype) followed by the norm"""
    _contig(h=h, y=y)
    hn = h + y
    return hn, rmsnorm(hn, weight, eps)


def rope_(x, cos, sin):
    """in place; the torch expression flatq_rope replaces (apply_rotary_pos_emb, every op rounded to x.dtype)"""
    _contig(x=x, cos=cos, sin=sin), _al16(x, 2, "rope x"), _al16(cos, 2, "rope cos"), _al16(sin, 2, "rope sin")
    _act_dtype(x, "rope x"), _act_dtype(cos, "rope cos"), _act_dtype(sin, "rope sin")
    heads, d = x.shape[-2], x.shape[-1]
    tokens = x.numel() // (heads * d)
    _req(d == 128 and cos.numel() == tokens * d and sin.numel() == tokens * d, "rope: shapes")
    xv, c, s = x.view(tokens, heads, d), cos.reshape(tokens, 1, d), sin.reshape(tokens, 1, d)
    rot = torch.cat((-xv[..., d // 2:], xv[..., :d // 2]), dim=-1)
    x.copy_((xv * c + rot * s).view(x.shape))
    return x


def silu_mul(gu):
    _contig(gu=gu), _al16(gu, 2, "silu_mul gu"), _act_dtype(gu, "silu_mul gu")
    inter = gu.shape[-1] // 2
    _req(inter % 8 == 0, "silu_mul: inter must be a multiple of 8")
    return torch.nn.functional.silu(gu[..., :inter]) * gu[..., inter:]


def hadamard(x, scale=1.0):
    _contig(x=x)
    n = x.shape[-1]
    _req(8 <= n <= 32768 and n & (n - 1) == 0, f"hadamard: n={n} must be a power of two in [8, 32768]")
    _req(x.numel() > 0, "hadamard: rows"), _al16(x, 2, "hadamard x"), _act_dtype(x, "hadamard x")
    return torch.from_numpy(O.hadamard_transform(x.float().numpy(), scale)).to(x.dtype)


def hadamard_kn(x, hadK, K, scale):
    _contig(x=x)
    n = x.shape[-1]
    _req(4 <= K <= 172 and K % 4 == 0 and n % K == 0, f"hadamard_kn: K={K} must be a multiple of 4 in [4, 172] dividing n={n}")
    n2 = n // K
    _req(n2 >= 8 and n2 & (n2 - 1) == 0 and n <= 32768, f"hadamard_kn: n/K={n2} must be a power of two >= 8, n <= 32768")
    _req((n + n // 32 + 2) * 4 + K * K * 2 <= 227 * 1024, "hadamard_kn: row does not fit shared memory")
    _req(x.numel() > 0, "hadamard_kn: rows"), _al16(x, 2, "hadamard_kn x"), _act_dtype(x, "hadamard_kn x")
    _req(tuple(hadK.shape) == (K, K), "hadamard_kn: hadK shape")
    y = O.matmul_hadU_kn(x.float().numpy(), hadK.float().numpy(), scale, fp16_intermediate=x.dtype == torch.float16)
    return torch.from_numpy(y).to(x.dtype)


def kv_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot, layer, want_k=False):
    """Stand-in of flatq_kv_fused built from the stand-ins of the launches it replaces (same bits by construction)."""
    hq, hkv, d = q.shape[-2], k.shape[-2], q.shape[-1]
    _req(d == 128, "kv_fused: head_dim"), _act_dtype(q, "kv_fused q")
    tokens = q.numel() // (hq * d)
    q2, k2, v2 = q.reshape(tokens, hq, d).contiguous(), k.reshape(tokens, hkv, d).contiguous(), v.reshape(tokens, hkv, d).contiguous()
    if cos_t is not None:
        c, s_ = cos_t[pos.long()].contiguous(), sin_t[pos.long()].contiguous()
        q2, k2 = rope_(q2.clone(), c, s_), rope_(k2.clone(), c, s_)
    if t_q is not None:
        q2 = kron_transform(q2.reshape(-1, d).contiguous(), None, t_q, 1, d).view(tokens, hq, d)
    if t_k is not None:
        k2 = kron_transform(k2.reshape(-1, d).contiguous(), None, t_k, 1, d).view(tokens, hkv, d)
    kv_quant(k2.contiguous(), kv_data, kv_param, slot, layer, 0)
    kv_quant(v2, kv_data, kv_param, slot, layer, 1)
    return q2, (k2 if want_k else None)


def kv_quant(x, kv_data=None, kv_param=None, slot=None, layer=0, kv_select=0):
    _contig(x=x, kv_data=kv_data, kv_param=kv_param, slot=slot)
    tokens, heads, d = x.shape
    _req(d == 128, f"kv_quant: head_dim={d} (only 128)")
    _req(tokens > 0 and heads > 0, "kv_quant: empty")
    _dtype(x, torch.float16, "kv_quant x")
    if kv_data is not None:
        _req(kv_param is not None and slot is not None and kv_select in (0, 1) and 0 <= layer < kv_data.shape[1],
             "kv_quant: paged-cache arguments")
        _dtype(kv_data, torch.uint8, "kv_data"), _dtype(kv_param, torch.float16, "kv_param"), _dtype(slot, torch.int32, "slot")
        _req(slot.numel() == tokens, "kv_quant: one slot per token")
        _req(kv_data.dim() == 6 and kv_data.shape[2] == 2 and kv_data.shape[3] == heads and kv_data.shape[5] == d // 2,
             f"kv_data layout {tuple(kv_data.shape)}")
        _req(tuple(kv_param.shape) == tuple(kv_data.shape[:5]) + (2,), f"kv_param layout {tuple(kv_param.shape)}")
        _req(int(slot.max()) < kv_data.shape[0] * kv_data.shape[4] and int(slot.min()) >= 0, "kv_quant: slot out of the page pool")
    CALLS.append(("kv_quant", (heads, kv_data is not None)))
    packed, s, z = O.kv_asym_quant(x)  # x fp16 [tokens, heads, d]: the reference's fp16 op-by-op arithmetic
    par = torch.cat([s, z], -1).to(torch.float16)
    if kv_data is not None:
        ps = kv_data.shape[4]
        page, off = (slot // ps).long(), (slot % ps).long()
        kv_data[page, layer, kv_select, :, off] = packed
        kv_param[page, layer, kv_select, :, off] = par
    return packed, par


def kv_decode_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot, kv_indptr, kv_indices, last_page_len, layer,
                    sm_scale=None, max_tokens_hint=0):
    """Stand-in of flatq_kv_decode_fused: the two launches it folds together (same bits by construction)."""
    hq, hkv = q.shape[-2], k.shape[-2]
    _req(q.numel() // (hq * 128) == q.shape[0] and 2 <= hq // hkv <= 16, "kv_decode_fused: one token per sequence, 2..16 query heads per kv head")
    _req(cos_t is not None and t_q is not None and t_k is not None, "kv_decode_fused: RoPE tables and both transforms")
    q_t, _ = kv_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, kv_data, kv_param, slot, layer)
    CALLS.append(("kv_decode_fused", (q.shape[0], hq, hkv)))
    return kv_decode_i4(q_t, kv_data, kv_param, kv_indptr, kv_indices, last_page_len, layer, sm_scale, None, max_tokens_hint)


def kv_decode_i4(q, kv_data, kv_param, kv_indptr, kv_indices, last_page_len, layer=0, sm_scale=None, split=None,
                 max_tokens_hint=0):
    _contig(q=q, kv_data=kv_data, kv_param=kv_param, kv_indptr=kv_indptr, kv_indices=kv_indices, last_page_len=last_page_len)
    b, hq, d = q.shape
    hkv, ps = kv_data.shape[3], kv_data.shape[4]
    _req(d == 128, f"kv_decode: head_dim={d} (only 128)")
    _req(b > 0 and hq > 0 and hkv > 0 and hq % hkv == 0, "kv_decode: heads")
    _act_dtype(q, "kv_decode q"), _dtype(kv_data, torch.uint8, "kv_data"), _dtype(kv_param, torch.float16, "kv_param")
    for name, t in (("kv_indptr", kv_indptr), ("kv_indices", kv_indices), ("last_page_len", last_page_len)):
        _dtype(t, torch.int32, name)
    _req(kv_indptr.numel() == b + 1 and last_page_len.numel() == b, "kv_decode: page-table sizes")
    _req(int(kv_indices.max()) < kv_data.shape[0], "kv_decode: page index out of the pool")
    _req(0 < int(last_page_len.min()) and int(last_page_len.max()) <= ps, "kv_decode: last_page_len in (0, page_size]")
    if max_tokens_hint:  # the bound the split-KV kernel sizes its grid from must cover every sequence
        n_max = max((int(kv_indptr[i + 1] - kv_indptr[i]) - 1) * ps + int(last_page_len[i]) for i in range(b))
        _req(max_tokens_hint >= n_max, f"kv_decode: max_tokens_hint {max_tokens_hint} < longest sequence {n_max}")
    CALLS.append(("kv_decode_i4", (hq, hkv, ps)))
    outs = []
    for i in range(b):
        pages = kv_indices[kv_indptr[i]:kv_indptr[i + 1]].long()
        n = (len(pages) - 1) * ps + int(last_page_len[i])
        gather = lambda t, sel: t[pages, layer, sel].permute(0, 2, 1, 3).reshape(len(pages) * ps, hkv, -1)[:n]  # noqa: E731
        rep = lambda t: t.repeat_interleave(hq // hkv, 1)[None]  # noqa: E731
        kp, vp = gather(kv_data, 0), gather(kv_data, 1)
        kpar, vpar = gather(kv_param, 0).float(), gather(kv_param, 1).float()
        o = O.decode_attention_i4(q[i:i + 1].float() * (1.0 if sm_scale is None else sm_scale * d ** 0.5),
                                  rep(kp), rep(kpar[..., :1]), rep(kpar[..., 1:]),
                                  rep(vp), rep(vpar[..., :1]), rep(vpar[..., 1:]))
        outs.append(o)
    return torch.cat(outs).to(q.dtype)


FUNCTIONS = ("kron_transform", "act_quant", "transform_quant", "quantize_rows", "gemm_w4a8", "gemm_w4a8_silu", "gemm_w4a8_grouped",
             "linear_fused", "linear_fused_silu", "pack_i4",
             "hadamard", "hadamard_kn", "kv_fused", "kv_quant", "kv_decode_i4", "kv_decode_fused", "rmsnorm", "add_rmsnorm", "rope_", "silu_mul")


def install(monkeypatch):
    """Route flatquant_b200.ops.<fn> to the CPU stand-ins for the duration of a test."""
    import sys

    from flatquant_b200 import ops
    me = sys.modules[__name__]
    for name in FUNCTIONS:
        monkeypatch.setattr(ops, name, getattr(me, name))

