"""Grouped (mixture-of-experts) W4A4 GEMM on the B200 against the verified per-expert launches: bit-identical outputs.

First hardware run: round 2 (10 passed, profiles/r02_first_gpu_call.md); `grouped=True` is the default of
flatquant_b200.moe.FlatQuantRoutedExperts since.  Also checked without a GPU: the tile schedule
(tests/test_grouped_sched.py, same header compiled for the host) and the host logic on the oracle-backed stand-ins
(tests/test_deepseek_block_host.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import flatq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _case(dev, counts, n, k, out_dtype=torch.float16, with_bias=False, hint=0, seed=0):
    from flatquant_b200 import ops
    g = torch.Generator(device=dev).manual_seed(seed)
    e, m = len(counts), int(sum(counts))
    xq8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev, generator=g)
    sx = torch.rand(m, device=dev, generator=g) * 0.05 + 0.01
    w = torch.randint(0, 256, (e, n, k // 2), dtype=torch.uint8, device=dev, generator=g)
    sw = torch.rand(e, n, device=dev, generator=g) * 0.02 + 0.01
    bias = torch.randn(e, n, device=dev, generator=g).to(out_dtype) if with_bias else None
    off = torch.zeros(e + 1, dtype=torch.int32, device=dev)
    off[1:] = torch.cumsum(torch.tensor(counts, device=dev), 0)
    y = torch.full((m, n), float("nan"), dtype=out_dtype, device=dev)
    ops.gemm_w4a8_grouped(xq8, sx, w, sw, off, bias, out_dtype, out=y, rows_per_group_hint=hint)
    want = torch.empty_like(y)
    r0 = 0
    for i, c in enumerate(counts):
        if c:
            ops.gemm_w4a8(xq8[r0:r0 + c], sx[r0:r0 + c], w[i], sw[i], None if bias is None else bias[i], out_dtype,
                          out=want[r0:r0 + c])
        r0 += c
    torch.cuda.synchronize()
    assert torch.equal(y, want), (counts, n, k, out_dtype, hint)
    return xq8, sx, w, sw, y


@pytest.mark.parametrize("counts,n,k,hint", [
    ([5], 256, 256, 0),                                   # one group, one tile
    ([0, 0, 7, 0], 100, 128, 0),                          # empty groups around one; N < 128 (weight rows of the next
                                                          # expert inside the TMA box must be masked)
    ([1] * 9, 130, 512, 0),                               # decode-like: one row per expert, N = 128 + 2
    ([37, 0, 64, 3, 129], 384, 1024, 0),                  # ragged, token tile from the mean (47 rows -> 128)
    ([37, 0, 64, 3, 129], 384, 1024, 4),                  # same with the smallest token tile (16)
    ([20, 9, 0, 31], 256, 256, 12),                       # 32-token tiles
    ([300, 0, 0, 1], 4096, 2048, 20),                     # one hot expert, several 64-token blocks, full-width N
    ([90, 250, 513], 512, 4096, 256),                     # 256-token single-CTA tiles
])
def test_grouped_gemm_equals_per_expert_launches(dev, counts, n, k, hint):
    xq8, sx, w, sw, y = _case(dev, counts, n, k, hint=hint)
    # and the oracle on the first non-empty group (K-permuted int8 operand format, see test_gpu_parity.py)
    e = next(i for i, c in enumerate(counts) if c)
    r0, c = int(sum(counts[:e])), counts[e]
    perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    c8 = xq8[r0:r0 + c].cpu().numpy().reshape(c, k // 8, 8)
    codes = np.empty_like(c8)
    codes[:, :, perm] = c8
    acc = O.gemm_i4(codes.reshape(c, k), O.unpack_i4(w[e].cpu().numpy()))
    ref = O.sym_dequant(acc, sx[r0:r0 + c].cpu().numpy(), sw[e].cpu().numpy()).astype(np.float16)
    assert np.array_equal(y[r0:r0 + c].cpu().numpy(), ref)


def test_grouped_gemm_bf16_bias_and_v3_shape(dev):
    _case(dev, [12, 0, 40, 7], 256, 512, out_dtype=torch.bfloat16, with_bias=True, seed=3)
    rng = np.random.default_rng(1)
    counts = np.bincount(rng.choice(64, 2048, p=rng.dirichlet(np.full(64, 0.3))), minlength=64).tolist()
    _case(dev, counts, 4096, 7168, seed=5)   # DeepSeek-V3 w1;w3 of 64 experts, skewed routing
    _case(dev, counts, 7168, 2048, seed=6)   # w2


def test_grouped_moe_equals_loop_on_gpu(dev):
    from flatquant_b200.moe import FlatQuantRoutedExperts
    torch.manual_seed(9)
    T_, dim, inter, E, top_k = 200, 1024, 512, 8, 2
    (a1, b1), (a2, b2) = O.get_decompose_dim(dim), O.get_decompose_dim(inter)
    mk = lambda n: (torch.eye(n) * 0.75 + torch.randn(n, n) * (0.5 / n ** 0.5)).half().to(dev)  # noqa: E731
    args = (torch.randint(0, 256, (E, 2 * inter, dim // 2), dtype=torch.uint8, device=dev),
            (torch.rand(E, 2 * inter) * 0.02 + 0.01).to(dev),
            torch.randint(0, 256, (E, dim, inter // 2), dtype=torch.uint8, device=dev),
            (torch.rand(E, dim) * 0.02 + 0.01).to(dev), (mk(a1), mk(b1)), (mk(a2), mk(b2)), (0.98, 0.98), (0.98, 0.98))
    x = torch.randn(T_, dim, device=dev).half()
    indices = torch.multinomial(torch.tensor([0.4, 0.25, 0.15, 0.1, 0.05, 0.05, 0.0, 0.0]).repeat(T_, 1), top_k).to(dev)
    weights = torch.rand(T_, top_k, device=dev).half()
    y_loop = FlatQuantRoutedExperts(*args, grouped=False)(x, weights, indices)
    y_grp = FlatQuantRoutedExperts(*args, grouped=True)(x, weights, indices)
    torch.cuda.synchronize()
    assert torch.equal(y_loop, y_grp)
