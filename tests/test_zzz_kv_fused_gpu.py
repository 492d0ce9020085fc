"""flatq_kv_fused (RoPE -> K / Q transform -> asymmetric 4-bit KV quantise -> page scatter in one launch,
reading the strided q / k / v views of the fused qkv GEMM output in place) against the launches it replaces -- flatq_rope,
flatq_kron_transform (right-only), flatq_kv_quant -- on the same inputs: BIT-IDENTICAL q_t, k_t, cache pages and params.
Those launches are themselves pinned to the reference (test_gpu_parity.py: kv_quant bit-exact vs the reference-generated
fixture, RoPE vs HF's expression in test_zzz_glue_kernels_gpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _unfused(ops, q, k, v, cos_t, sin_t, pos, t_q, t_k, pages, params, slot, layer):
    tokens = pos.numel()
    hq, hkv = q.shape[-2], k.shape[-2]
    q2, k2, v2 = (t.reshape(tokens, -1, 128).clone() for t in (q, k, v))   # rope_ works in place: never on the caller's tensors
    if cos_t is not None:
        c, s = cos_t[pos.long()].contiguous(), sin_t[pos.long()].contiguous()
        q2, k2 = ops.rope_(q2, c, s), ops.rope_(k2, c, s)
    if t_q is not None:
        q2 = ops.kron_transform(q2.view(-1, 128), None, t_q, 1, 128).view(tokens, hq, 128)
        k2 = ops.kron_transform(k2.view(-1, 128), None, t_k, 1, 128).view(tokens, hkv, 128)
    ops.kv_quant(k2.contiguous(), pages, params, slot, layer, 0)
    ops.kv_quant(v2, pages, params, slot, layer, 1)
    return q2, k2


@pytest.mark.parametrize("b,s,hq,hkv,fused_view,rope,trans", [
    (1, 2048, 32, 8, True, True, True),      # Llama-3-8B prefill, views of one qkv buffer
    (2, 37, 32, 8, True, True, True),        # ragged row count (tail CTAs, tail warps)
    (16, 1, 32, 8, True, True, True),        # decode bsz 16
    (1, 1, 32, 8, True, True, True),         # decode bsz 1
    (3, 5, 4, 2, False, True, True),         # separate contiguous q / k / v, small GQA
    (2, 9, 8, 8, True, False, True),         # no RoPE
    (2, 9, 8, 1, True, True, False),         # no transform (trans == "none"), kv heads replicated under TP
])
def test_kv_fused_equals_the_launches_it_replaces(dev, b, s, hq, hkv, fused_view, rope, trans):
    from flatquant_b200 import ops
    torch.manual_seed(b * 1000 + s + hq)
    tokens, D, L, layer, ps = b * s, 128, 2, 1, 64
    if fused_view:
        qkv = torch.randn(b, s, (hq + 2 * hkv) * D, device=dev).half()
        q, k, v = (t.reshape(b, s, -1, D) for t in torch.split(qkv, [hq * D, hkv * D, hkv * D], dim=-1))
    else:
        q, k, v = (torch.randn(b, s, h, D, device=dev).half() for h in (hq, hkv, hkv))
    P = 4096
    ang = torch.arange(P, device=dev)[:, None].float() * (1.0 / 500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))[None]
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = (emb.cos().half(), emb.sin().half()) if rope else (None, None)
    pos = (torch.arange(s, device=dev) + 1000).repeat(b).int()
    t_k = (torch.eye(D, device=dev) * 0.75 + torch.randn(D, D, device=dev) * (0.5 / D ** 0.5)).half() if trans else None
    t_q = torch.linalg.inv(t_k.double()).t().float().half().contiguous() if trans else None
    n_pages = b * ((s + ps - 1) // ps)
    slot = (torch.arange(b, device=dev)[:, None] * ((s + ps - 1) // ps) * ps + torch.arange(s, device=dev)[None]).reshape(-1).int()
    caches = []
    for _ in range(2):
        caches.append((torch.zeros(n_pages, L, 2, hkv, ps, D // 2, dtype=torch.uint8, device=dev),
                       torch.zeros(n_pages, L, 2, hkv, ps, 2, dtype=torch.float16, device=dev)))
    q_ref, k_ref = _unfused(ops, q, k, v, cos_t, sin_t, pos, t_q, t_k, caches[0][0], caches[0][1], slot, layer)
    q_t, k_t = ops.kv_fused(q, k, v, cos_t, sin_t, pos if rope else None, t_q, t_k, caches[1][0], caches[1][1], slot, layer, want_k=True)
    torch.cuda.synchronize()
    assert q_t.shape == (tokens, hq, D) and k_t.shape == (tokens, hkv, D)
    assert torch.equal(q_t, q_ref.view(tokens, hq, D)), "q_t"
    assert torch.equal(k_t, k_ref.view(tokens, hkv, D)), "k_t"
    assert torch.equal(caches[0][0], caches[1][0]), "packed pages"
    assert torch.equal(caches[0][1], caches[1][1]), "scale / zero params"
    assert caches[1][0][:, 0].count_nonzero() == 0   # the other layer is untouched


@pytest.mark.parametrize("bsz,hq,hkv,ctx", [(1, 32, 8, 2048), (16, 32, 8, 700), (3, 64, 8, 129), (2, 16, 8, 5), (4, 28, 4, 1), (1, 40, 8, 4097)])
def test_decode_step_with_the_kv_front_folded_into_the_attention_kernel(dev, bsz, hq, hkv, ctx):
    """flatq_kv_decode_fused == flatq_kv_fused followed by flatq_kv_decode_i4_split, bit for bit: the packed pages and the
    (scale, zero) params of the new token (written by the CTA whose split holds it), and the attention output (q_t is
    produced by the same mma sequence, the rest is the same kernel).  Contexts shorter than the split count (empty splits),
    a first token (ctx = 1), strided q / k / v views of a fused qkv output, several sequences with their own positions."""
    from flatquant_b200 import ops
    D, L, layer = 128, 2, 1
    torch.manual_seed(bsz * 1000 + hq + ctx)
    ps = ctx + 3
    qkv = torch.randn(bsz, (hq + 2 * hkv) * D, device=dev).half()
    q, k, v = (t.reshape(bsz, 1, -1, D) for t in torch.split(qkv, [hq * D, hkv * D, hkv * D], dim=-1))
    P = ctx + 8
    ang = torch.arange(P, device=dev)[:, None].float() * (1.0 / 500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))[None]
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = emb.cos().half(), emb.sin().half()
    pos = torch.full((bsz,), ctx - 1, device=dev).int()
    t_k = (torch.eye(D, device=dev) * 0.75 + torch.randn(D, D, device=dev) * (0.5 / D ** 0.5)).half()
    t_q = torch.linalg.inv(t_k.double()).t().float().half().contiguous()
    pages0 = torch.randint(0, 256, (bsz, L, 2, hkv, ps, D // 2), dtype=torch.uint8, device=dev)
    params0 = torch.stack((torch.rand(bsz, L, 2, hkv, ps, device=dev) * 0.2 + 0.05, torch.rand(bsz, L, 2, hkv, ps, device=dev) * 1.5), -1).half()
    slot = (torch.arange(bsz, device=dev) * ps + (ctx - 1)).int()
    indptr = torch.arange(bsz + 1, dtype=torch.int32, device=dev)
    indices = torch.arange(bsz, dtype=torch.int32, device=dev)
    last = torch.full((bsz,), ctx, dtype=torch.int32, device=dev)
    assert ops.kv_decode_fused_ok(q, k, cos_t, t_q, t_k)
    pages_a, params_a = pages0.clone(), params0.clone()
    q_t, _ = ops.kv_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, pages_a, params_a, slot, layer)
    want = ops.kv_decode_i4(q_t, pages_a, params_a, indptr, indices, last, layer=layer, max_tokens_hint=ps)
    pages_b, params_b = pages0.clone(), params0.clone()
    got = ops.kv_decode_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, pages_b, params_b, slot, indptr, indices, last, layer=layer,
                              max_tokens_hint=ps)
    torch.cuda.synchronize()
    assert torch.equal(pages_a, pages_b) and torch.equal(params_a, params_b)
    assert not torch.equal(pages_b, pages0)          # the new token was written
    assert torch.equal(got, want)
    assert bool(torch.isfinite(got.float()).all()) and float(got.float().abs().max()) > 0
