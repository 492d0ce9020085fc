"""One-pass glue kernels (csrc/glue.cuh: flatq_rmsnorm, flatq_rope, flatq_silu_mul) against the torch expressions they
replace, on the GPU in fp16.  First hardware run: round 2 (10 passed, profiles/r02_first_gpu_call.md);
the fused glue is the default of the model stack since (FLATQ_FUSED_GLUE=0 opts out)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("rows,k", [(5, 512), (2048, 4096), (33, 8192), (7, 14336)])
def test_rmsnorm_matches_hf_expression(dev, rows, k):
    from flatquant_b200 import ops
    torch.manual_seed(rows + k)
    x = (torch.randn(rows, k, device=dev) * 3).half()
    w = (torch.rand(k, device=dev) + 0.5).half()
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).half()
    got = ops.rmsnorm(x, w, 1e-5)
    # the fp32 sum of squares is accumulated in another order: the normalised value may round differently by one fp16 ulp
    d = (got.float() - ref.float()).abs()
    assert float((d / ref.float().abs().clamp(min=1e-3)).max()) < 2.0 ** -9 and float((d > 0).float().mean()) < 0.02
    bare = (xf * torch.rsqrt(xf.pow(2).sum(-1, keepdim=True) / k + 1e-5)).half()   # the weight-less norm (deploy.nn.RMSNorm)
    d = (ops.rmsnorm(x, None, 1e-5, mean_dim=k).float() - bare.float()).abs()
    assert float((d / bare.float().abs().clamp(min=1e-3)).max()) < 2.0 ** -9


@pytest.mark.parametrize("rows,k", [(1, 4096), (2048, 4096), (16, 4096), (5, 14336)])
def test_add_rmsnorm_equals_add_then_rmsnorm(dev, rows, k):
    """The residual add folded into the norm: h_out is bit-identical to torch's fp16 `h + y`, and the normalised output
    is bit-identical to flatq_rmsnorm of that sum (same kernel body, same summation order)."""
    from flatquant_b200 import ops
    torch.manual_seed(rows * 7 + k)
    h = (torch.randn(rows, k, device=dev) * 2).half()
    y = torch.randn(rows, k, device=dev).half()
    w = (torch.rand(k, device=dev) + 0.5).half()
    h_out, normed = ops.add_rmsnorm(h, y, w, 1e-5)
    assert torch.equal(h_out, h + y)
    assert torch.equal(normed, ops.rmsnorm(h + y, w, 1e-5))


@pytest.mark.parametrize("b,s,h", [(1, 7, 4), (2, 2048, 32), (16, 1, 8)])
def test_rope_is_bit_identical_to_the_torch_expression(dev, b, s, h):
    from flatquant_b200 import ops
    from flatquant_b200.deploy.transformers import modeling_llama as ML
    torch.manual_seed(b + s + h)
    x = torch.randn(b, s, h, 128, device=dev).half()
    inv = 1.0 / (500000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32, device=dev) / 128))
    pos = torch.arange(3, 3 + s, device=dev)[None].expand(b, s)
    cos, sin = ML.rope_cos_sin(inv, pos, torch.float16)
    half = 64
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    ref = x * cos + rot * sin
    got = ops.rope_(x.clone(), cos.reshape(b * s, 128).contiguous(), sin.reshape(b * s, 128).contiguous())
    assert torch.equal(got, ref)   # same three fp16 roundings per element


@pytest.mark.parametrize("rows,inter", [(3, 64), (2048, 14336), (17, 1792)])
def test_silu_mul_matches_the_torch_expression(dev, rows, inter):
    from flatquant_b200 import ops
    torch.manual_seed(rows + inter)
    gu = (torch.randn(rows, 2 * inter, device=dev) * 2).half()
    ref = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
    got = ops.silu_mul(gu)
    d = (got.float() - ref.float()).abs()
    # torch's silu kernel and expf may differ in the last fp32 bit before the fp16 rounding: at most one fp16 ulp, rarely
    assert float((d / ref.float().abs().clamp(min=1e-3)).max()) < 2.0 ** -9 and float((d > 0).float().mean()) < 1e-3


@pytest.mark.parametrize("m,inter,k", [(1, 512, 512), (16, 1032, 512), (100, 1024, 1024), (129, 1032, 512), (300, 520, 2048),
                                       (2048, 14336, 4096), (1, 14336, 4096), (2, 1032, 4096), (2, 520, 5120), (1, 1024, 8192),
                                       (2, 14336, 4096), (1, 72, 7168), (8, 14336, 4096), (16, 14336, 4096), (32, 14336, 4096),
                                       (16, 14344, 4096)])
def test_gate_up_gemm_with_silu_epilogue_is_bit_identical_to_gemm_then_silu_mul(dev, m, inter, k):
    """flatq_gemm_w4a8_silu on row-interleaved [gate; up] weights == flatq_gemm_w4a8 on the concatenated weights followed
    by flatq_silu_mul, bit for bit (same fp16 rounding of the GEMM output, same fp32 SiLU, same fp16 product): the
    epilogue-fused leg (M > 128, CTA-pair kernel, ragged last tiles included), the decode leg (M <= 2: the GEMV's warps own
    (gate, up) row pairs) and the workspace leg (split-K, or a GEMV width without a pair instantiation)."""
    from flatquant_b200 import ops
    torch.manual_seed(m + inter + k)
    xq8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev)
    sx = torch.rand(m, device=dev) * 0.05 + 0.02
    wg, wu = (torch.randint(-8, 8, (inter, k), dtype=torch.int8, device=dev) for _ in range(2))
    sg, su = (torch.rand(inter, device=dev) * 0.05 + 0.02 for _ in range(2))
    w_cat, s_cat = ops.pack_i4(torch.cat([wg, wu], 0)), torch.cat([sg, su]).contiguous()
    w_il = ops.pack_i4(torch.stack([wg, wu], 1).reshape(2 * inter, k).contiguous())
    s_il = torch.stack([sg, su], 1).reshape(-1).contiguous()
    ref = ops.silu_mul(ops.gemm_w4a8(xq8, sx, w_cat, s_cat, None, torch.float16))
    got = ops.gemm_w4a8_silu(xq8, sx, w_il, s_il)
    got2 = ops.gemm_w4a8_silu(xq8, sx, w_il, s_il)
    torch.cuda.synchronize()
    assert got.shape == (m, inter) and torch.equal(got, ref) and torch.equal(got, got2)
    assert float(ref.float().abs().mean()) > 1e-3   # not a comparison of zeros
