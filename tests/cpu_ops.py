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


def _h(t):
    return t.half().float() if EMULATE_FP16 else t.float()


def _transform(x2, left, right, n1, n2):
    xm = _h(x2).reshape(-1, n1, n2)
    if right is not None:
        xm = _h(torch.matmul(xm, _h(right)))
    if left is not None:
        xm = _h(torch.matmul(_h(left).t(), xm))
    return xm.reshape(x2.shape[0], n1 * n2)


def kron_transform(x, left, right, n1, n2):
    if left is None:
        n1 = 1
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
    z = _transform(x.reshape(-1, k), left, right, n1, n2)
    q, s = _scale_codes(z, clip_max, clip_min)
    if out is not None:
        out[0].copy_(q)
        out[1].copy_(s)
        return out
    return q, s


def transform_quant(x, left, right, n1, n2, clip_max=1.0, clip_min=1.0):
    q, s = act_quant(x, left, right, n1, n2, clip_max, clip_min)
    return torch.from_numpy(O.pack_i4(q.numpy())), s.half(), s


def quantize_rows(x, clip_max=1.0, clip_min=1.0, half_math=True):
    q, s = _scale_codes(x.reshape(-1, x.shape[-1]).float(), clip_max, clip_min)
    return torch.from_numpy(O.pack_i4(q.numpy())), s.half(), s


def gemm_w4a8(xq8, s_x, w_packed, s_w, bias=None, out_dtype=torch.float16, out=None):
    wq = O.unpack_i4(w_packed.numpy())
    acc = O.gemm_i4(xq8.numpy(), wq)
    y = O.sym_dequant(acc, s_x.numpy(), s_w.float().numpy(), None if bias is None else bias.float().numpy())
    y = torch.from_numpy(y).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def linear_fused(x, left, right, n1, n2, clip_max, clip_min, w_packed, s_w, bias=None, out_dtype=torch.float16):
    k = n1 * n2
    q, s = act_quant(x.reshape(-1, k), left, right, n1, n2, clip_max, clip_min)
    return gemm_w4a8(q, s, w_packed, s_w, bias, out_dtype).reshape(x.shape[:-1] + (w_packed.shape[0],))


def pack_i4(q):
    return torch.from_numpy(O.pack_i4(q.numpy()))


def hadamard(x, scale=1.0):
    return torch.from_numpy(O.hadamard_transform(x.float().numpy(), scale)).to(x.dtype)


def kv_quant(x, kv_data=None, kv_param=None, slot=None, layer=0, kv_select=0):
    packed, s, z = O.kv_asym_quant(x)  # x fp16 [tokens, heads, d]: the reference's fp16 op-by-op arithmetic
    par = torch.cat([s, z], -1).to(torch.float16)
    if kv_data is not None:
        ps = kv_data.shape[4]
        page, off = (slot // ps).long(), (slot % ps).long()
        kv_data[page, layer, kv_select, :, off] = packed
        kv_param[page, layer, kv_select, :, off] = par
    return packed, par


def kv_decode_i4(q, kv_data, kv_param, kv_indptr, kv_indices, last_page_len, layer=0, sm_scale=None):
    b, hq, d = q.shape
    hkv, ps = kv_data.shape[3], kv_data.shape[4]
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


FUNCTIONS = ("kron_transform", "act_quant", "transform_quant", "quantize_rows", "gemm_w4a8", "linear_fused", "pack_i4",
             "hadamard", "kv_quant", "kv_decode_i4")


def install(monkeypatch):
    """Route flatquant_b200.ops.<fn> to the CPU stand-ins for the duration of a test."""
    import sys

    from flatquant_b200 import ops
    me = sys.modules[__name__]
    for name in FUNCTIONS:
        monkeypatch.setattr(ops, name, getattr(me, name))

