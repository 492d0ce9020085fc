"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures
generated from the reference.  Bit-exact for pack/unpack indexing, codes given identical inputs, and
int32 accumulators; fp16 tolerances are stated where used."""
import numpy as np
import pytest
import torch

from oracle import flatq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from flatquant_b200 import ops as _ops
    return _ops


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ------------------------------------------------------------------ a8: storage format, bit exact
def test_pack_unpack_golden(ops, dev, golden):
    g = golden("pack_i4")
    assert np.array_equal(ops.pack_i4(T(g["q"]).to(dev)).cpu().numpy(), g["packed"])
    assert np.array_equal(ops.unpack_i4(T(g["packed"]).to(dev)).cpu().numpy(), g["unpacked"])


def test_pack_unpack_roundtrip_large(ops, dev):
    q = torch.randint(-8, 8, (2048, 4096), dtype=torch.int8, device=dev)
    p = ops.pack_i4(q)
    assert torch.equal(ops.unpack_i4(p), q.int())
    assert np.array_equal(p[:64].cpu().numpy(), O.pack_i4(q[:64].cpu().numpy()))


# ------------------------------------------------------------------ a10 / a12
def test_sym_quant_golden_and_oracle(ops, dev, golden):
    g = golden("sym_quant_half")
    got = ops.unpack_i4(ops.sym_quant(T(g["x"]).to(dev), T(g["scale"]).reshape(-1).to(dev))).cpu().numpy()
    assert np.array_equal(got, g["q"].astype(np.int32))
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((257, 1024)) * 5).astype(np.float16)
    s = (np.abs(x).max(1) / 7).astype(np.float16)
    got = ops.sym_quant(T(x).to(dev), T(s).to(dev)).cpu().numpy()
    assert np.array_equal(got, O.pack_i4(O.sym_quant_half(x, s)))


def test_sym_dequant_modes(ops, dev):
    rng = np.random.default_rng(1)
    acc = rng.integers(-900000, 900000, (96, 256)).astype(np.int32)
    sr = rng.random(96).astype(np.float16)
    sc = (rng.random(256) * 0.01).astype(np.float16)
    y = ops.sym_dequant(T(acc).to(dev), T(sr).to(dev), T(sc).to(dev)).cpu().numpy()
    assert np.array_equal(y, O.sym_dequant(acc, sr, sc).astype(np.float16))
    yk = ops.sym_dequant(T(acc).to(dev), T(sr).to(dev), T(sc).to(dev), ref_kernel_arith=True).cpu().numpy()
    assert np.array_equal(yk, O.sym_dequant_ref_kernel(acc, sr, sc))


# ------------------------------------------------------------------ a11: INT4 x INT4 -> INT32, bit exact
@pytest.mark.parametrize("m,n,k", [(1, 128, 64), (16, 128, 128), (100, 200, 192), (128, 256, 4096),
                                   (33, 1024, 14336), (2048, 4096, 4096), (300, 6144, 4096)])
def test_gemm_i4_exact(ops, dev, m, n, k):
    rng = np.random.default_rng(m * 7 + n)
    a = rng.integers(-8, 8, (m, k)).astype(np.int8)
    b = rng.integers(-8, 8, (n, k)).astype(np.int8)
    a[0, :] = -8
    b[0, :] = -8  # extreme accumulator: 64*K
    c = ops.gemm_i4(T(O.pack_i4(a)).to(dev), T(O.pack_i4(b)).to(dev)).cpu().numpy()
    assert np.array_equal(c, O.gemm_i4(a, b))


def test_linear4bit_matches_exact_dequant(ops, dev):
    rng = np.random.default_rng(5)
    m, n, k = 77, 384, 512
    a = rng.integers(-8, 8, (m, k)).astype(np.int8)
    b = rng.integers(-8, 8, (n, k)).astype(np.int8)
    sa = rng.random(m).astype(np.float16)
    sb = (rng.random(n) * 0.02).astype(np.float16)
    bias = rng.standard_normal(n).astype(np.float16)
    y = ops.linear4bit(T(O.pack_i4(a)).to(dev), T(sa).to(dev), T(O.pack_i4(b)).to(dev), T(sb).to(dev), T(bias).to(dev))
    ref = O.sym_dequant(O.gemm_i4(a, b), sa, sb, bias).astype(np.float16)
    assert np.array_equal(y.cpu().numpy(), ref)


# ------------------------------------------------------------------ a2 / a3: online transform
@pytest.mark.parametrize("dims", ["16x16", "14x16", "64x64"])
def test_kron_transform_golden(ops, dev, golden, dims):
    g = golden(f"kron_{dims}")
    n1, n2 = g["left"].shape[0], g["right"].shape[0]
    x = T(g["x"]).half()
    L, R = T(g["left"]).half(), T(g["right"]).half()
    z = ops.kron_transform(x.to(dev), L.to(dev), R.to(dev), n1, n2).float().cpu()
    ref = O.kronecker_matmul(x.float(), L.float(), R.float())  # same fp16-rounded inputs, fp32 math
    # tolerance: two fp16 roundings (intermediate + output) of values up to max|z|
    tol = 3 * 2.0 ** -11 * float(ref.abs().max()) * (n1 ** 0.5)
    assert (z - ref).abs().max() <= tol


def _emulated_codes(x, L, R, cmax, cmin, order):
    """CPU model of the kernels' arithmetic: fp32-accumulated products with the fp16 roundings the kernels apply
    (order 'xr' = L^T.(X.R) of the tcgen05 kernel / reference; 'lx' = (L^T.X).R of the HMMA kernel)."""
    n1, n2 = L.shape[0], R.shape[0]
    xm = x.float().reshape(-1, n1, n2)
    if order == "xr":
        y = torch.matmul(xm, R.float()).half().float()
        z = torch.matmul(L.float().t(), y).half().float()
    else:
        u = torch.matmul(L.float().t(), xm).half().float()
        z = torch.matmul(u, R.float()).half().float()
    z = z.reshape(x.shape[0], -1)
    hi = z.amax(1).clamp(min=0) * torch.tensor(cmax)
    lo = z.amin(1).clamp(max=0) * torch.tensor(cmin)
    m = torch.maximum(lo.abs(), hi)
    s = torch.where(m == 0, torch.ones_like(m), m / 7)
    q = torch.clamp(torch.round(z / s[:, None]), -8, 7).numpy().astype(np.int32)
    return q, s


# tcgen05 kernel, second design (kron_quant_ts.cuh, left factor first like the deployed reference):
#   (64,64) (112,128) (56,64) (32,64) (64,128) (128,128);  first design (kron_quant_tc.cuh, right first): (86,128) (48,64);
# HMMA kernel (left first): the rest
TS_SHAPES = {(64, 64), (112, 128), (56, 64), (32, 64), (64, 128), (128, 128)}


@pytest.mark.parametrize("n1,n2", [(64, 64), (112, 128), (86, 128), (56, 64), (32, 64), (64, 128), (128, 128), (48, 64),
                                   (32, 48), (128, 224), (8, 16), (14, 16)])
def test_transform_quant_codes(ops, dev, n1, n2):
    torch.manual_seed(n1 * 1000 + n2)
    tokens, k = 37, n1 * n2
    x = torch.randn(tokens, k).half()
    x[3] = 0  # all-zero token: scale must be 1, codes 0
    L = (torch.randn(n1, n1) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2) / n2 ** 0.5).half()
    cmax, cmin = 0.982, 0.95
    xq, s16, s32 = ops.transform_quant(x.to(dev), L.to(dev), R.to(dev), n1, n2, cmax, cmin)
    codes = O.unpack_i4(xq.cpu().numpy())
    assert s32[3].item() == 1.0 and not codes[3].any()
    assert torch.equal(s16.cpu(), s32.cpu().half())
    tc = n2 in (64, 128) and 32 <= n1 <= 128
    # (1) against a CPU model with the kernel's own rounding points.  Only the fp32 summation order inside the tensor
    #     core differs, which can move an fp16 rounding (and hence a code) in rare cases.
    q_em, s_em = _emulated_codes(x, L, R, cmax, cmin, "xr" if tc and (n1, n2) not in TS_SHAPES else "lx")
    d = np.abs(codes - q_em)
    assert d.max() <= 1 and (d != 0).mean() < 2e-4   # measured on the B200: <= 3e-5 (128,128), 0 on most shapes
    assert torch.allclose(s32.cpu(), s_em, rtol=1e-3)
    if not tc:
        # HMMA kernel: the quantiser is bit exact given the kernel's own fp16 transform output
        zk = ops.kron_transform(x.to(dev), L.to(dev), R.to(dev), n1, n2).cpu().float()
        hi = zk.amax(1).clamp(min=0) * torch.tensor(cmax)
        lo = zk.amin(1).clamp(max=0) * torch.tensor(cmin)
        m = torch.maximum(lo.abs(), hi)
        s_ref = torch.where(m == 0, torch.ones_like(m), m / 7)
        assert torch.equal(s32.cpu(), s_ref)
        q_ref = torch.clamp(torch.round(zk / s_ref[:, None]), -8, 7).numpy().astype(np.int32)
        assert np.array_equal(codes, q_ref)
    # (2) against the pure fp32 oracle: codes agree except for +-1 flips at rounding boundaries
    z32 = O.kronecker_matmul(x.float(), L.float(), R.float())
    hi, lo = z32.amax(1).clamp(min=0) * cmax, z32.amin(1).clamp(max=0) * cmin
    m = torch.maximum(lo.abs(), hi)
    so = torch.where(m == 0, torch.ones_like(m), m / 7)
    qo = torch.clamp(torch.round(z32 / so[:, None]), -8, 7).numpy().astype(np.int32)
    diff = np.abs(codes - qo)
    print(f"({n1},{n2}): flips vs fp32 oracle {(diff != 0).mean():.5f}, vs rounding-point model {(d != 0).mean():.5f}")
    assert diff.max() <= 1
    # measured on the B200: 0.04-0.06 % on the Llama shapes, 0.11 % at (8,16); bound = 2x the worst (SURVEY 8c proposed
    # >= 99.9 % identical for the deployed shapes: met)
    assert (diff != 0).mean() < (0.0025 if n1 * n2 < 1024 else 0.0012)
    assert torch.allclose(s32.cpu(), so, rtol=2e-3)
    # (3) the GEMM operand format (int8, K-permuted) carries the same codes
    xq8, sx = ops.act_quant(x.to(dev), L.to(dev), R.to(dev), n1, n2, cmax, cmin)
    perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    unperm = xq8.cpu().numpy().reshape(tokens, k // 8, 8)
    restored = np.empty_like(unperm)
    restored[:, :, perm] = unperm
    d8 = np.abs(restored.reshape(tokens, k).astype(np.int32) - codes)
    assert d8.max() <= (1 if tc else 0) and (d8 != 0).mean() < 2e-4
    assert torch.allclose(sx.cpu(), s32.cpu(), rtol=1e-3)


@pytest.mark.parametrize("heads", [16, 32, 64])
def test_left_only_transform_on_the_tcgen05_kernel(ops, dev, heads):
    """o_proj's online transform Z = O^T . A over the heads axis (A = attention output [heads, 128] per token): served by
    the two-factor tcgen05 kernel with an identity right factor since round 2 -- exact, so the codes must equal the
    left-only rounding model (one fp16 rounding after O^T . A) up to tensor-core summation order, and the legacy-MMA
    kernel's codes (FLATQ_KRON_HMMA semantics) on the same input."""
    torch.manual_seed(heads)
    tokens, d = 77, 128
    x = torch.randn(tokens, heads * d).half()
    L = (torch.eye(heads) * 0.8 + torch.randn(heads, heads) / heads ** 0.5 * 0.5).half()
    cmax, cmin = 0.982, 0.95
    xq, s16, s32 = ops.transform_quant(x.to(dev), L.to(dev), None, heads, d, cmax, cmin)
    codes = O.unpack_i4(xq.cpu().numpy())
    z = torch.matmul(L.float().t(), x.float().reshape(tokens, heads, d)).half().float().reshape(tokens, -1)
    hi, lo = z.amax(1).clamp(min=0) * torch.tensor(cmax), z.amin(1).clamp(max=0) * torch.tensor(cmin)
    m = torch.maximum(lo.abs(), hi)
    s = torch.where(m == 0, torch.ones_like(m), m / 7)
    q = torch.clamp(torch.round(z / s[:, None]), -8, 7).numpy().astype(np.int32)
    d_ = np.abs(codes - q)
    assert d_.max() <= 1 and (d_ != 0).mean() < 2e-4
    assert torch.allclose(s32.cpu(), s, rtol=1e-3)
    xq8, sx = ops.act_quant(x.to(dev), L.to(dev), None, heads, d, cmax, cmin)
    perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    un = xq8.cpu().numpy().reshape(tokens, heads * d // 8, 8)
    rest = np.empty_like(un)
    rest[:, :, perm] = un
    assert np.array_equal(rest.reshape(tokens, -1).astype(np.int32), codes)


def test_quantize_rows_deploy_semantics(ops, dev):
    """deploy.nn.Quantizer: fp16 scale = max|x|/7, codes = sym_quant(x, scale)."""
    torch.manual_seed(3)
    x = (torch.randn(65, 4096) * 2).half()
    xq, s16, _ = ops.quantize_rows(x.to(dev), 1.0, 1.0, half_math=True)
    s_ref = (x.abs().amax(1).float() / 7).half()
    assert torch.equal(s16.cpu(), s_ref)
    assert np.array_equal(xq.cpu().numpy(), O.pack_i4(O.sym_quant_half(x.numpy(), s_ref.numpy())))


# ------------------------------------------------------------------ the whole path vs the golden fixtures
@pytest.mark.parametrize("case", ["flatlinear_256_svd", "flatlinear_224_inv_bias", "flatlinear_4096_svd"])
def test_fused_linear_golden(ops, dev, golden, case):
    g = golden(case)
    n1, n2 = g["left"].shape[0], g["right"].shape[0]
    x = T(g["x"]).half().reshape(-1, n1 * n2)
    L, R = T(g["left"]).half(), T(g["right"]).half()
    cmax = float(torch.sigmoid(torch.tensor(float(g["ca_max"]))))
    cmin = float(torch.sigmoid(torch.tensor(float(g["ca_min"]))))
    bias = T(g["bias"]).half().to(dev) if g["bias"].size else None
    y = ops.linear_fused(x.to(dev), L.to(dev), R.to(dev), n1, n2, cmax, cmin, T(g["w_packed"]).to(dev),
                         T(g["s_w"]).reshape(-1).float().to(dev), bias).float().cpu().numpy()
    y_ref = g["y"].reshape(y.shape)  # the reference's fake-quant output (fp32 CPU)
    # fp16 tolerance (north star): inputs/factors rounded to fp16 and an fp16 transform move a few codes by one
    # step; each flip changes y by <= s_x*s_w*|w|.  Stated bound: 3% of max|y| worst case, 0.45% on average (measured on
    # the B200: max 2.3-2.8 %, mean 0.19-0.35 %; the fixtures' x is fp32 and is rounded to fp16 here).
    err = np.abs(y - y_ref)
    print(f"{case}: max err {err.max() / np.abs(y_ref).max():.5f}, mean err {err.mean() / np.abs(y_ref).max():.6f} of max|y|")
    assert err.max() <= 0.03 * np.abs(y_ref).max()
    assert err.mean() <= 0.0045 * np.abs(y_ref).max()
    # tight check: oracle evaluated on the same fp16-rounded inputs with the kernel's own codes is exact to fp16
    xq, _, s32 = ops.transform_quant(x.to(dev), L.to(dev), R.to(dev), n1, n2, cmax, cmin)
    acc = O.gemm_i4(O.unpack_i4(xq.cpu().numpy()), g["w_codes"])
    y2 = O.sym_dequant(acc, s32.cpu().numpy(), g["s_w"].reshape(-1),
                       g["bias"].astype(np.float16).astype(np.float32) if g["bias"].size else None)
    assert np.array_equal(y.astype(np.float16), y2.astype(np.float16))


def test_fused_linear_full_size_consistency(ops, dev):
    """BASELINE size (2048, 4096 -> 4096): fused output == oracle arithmetic on the kernel's own codes for
    sampled rows, and linear in the weight scale."""
    torch.manual_seed(11)
    m, n, n1, n2 = 2048, 4096, 64, 64
    k = n1 * n2
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / 8).half()
    R = (torch.randn(n2, n2, device=dev) / 8).half()
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    wp = ops.pack_i4(wq)
    s_w = torch.rand(n, device=dev) * 0.02 + 0.01
    y = ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, wp, s_w)
    y2 = ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, wp, s_w * 2)
    assert torch.equal((y.float() * 2).half(), y2)  # power-of-two scaling is exact
    xq, _, s32 = ops.transform_quant(x, L, R, n1, n2, 0.98, 0.98)
    rows = torch.tensor([0, 1, 127, 128, 1000, 2047])
    codes = O.unpack_i4(xq[rows.to(dev)].cpu().numpy())
    acc = O.gemm_i4(codes, wq.cpu().numpy())
    ref = O.sym_dequant(acc, s32[rows.to(dev)].cpu().numpy(), s_w.cpu().numpy()).astype(np.float16)
    assert np.array_equal(y[rows.to(dev)].cpu().numpy(), ref)


# ------------------------------------------------------------------ reference operator API (drop-in)
def _args():
    import types
    return types.SimpleNamespace(w_bits=4, a_bits=4, w_asym=False, a_asym=False, lac=True, lwc=True, a_groupsize=-1)


def test_flat_quantized_linear_dropin(dev, golden):
    """The reference call pattern: to_eval_mode -> reparameterize(qa_trans) -> RTN -> fl(trans(x))."""
    from flatquant_b200.flat_linear import FlatQuantizedLinear
    from flatquant_b200.quant_utils import WeightQuantizer
    from flatquant_b200.trans_utils import SVDDecomposeTransMatrix
    g = golden("flatlinear_256_svd")
    n1, n2 = g["left"].shape[0], g["right"].shape[0]
    n = g["w0"].shape[0]
    lin = torch.nn.Linear(n1 * n2, n, bias=False)
    lin.weight.data = T(g["w0"]).clone()
    tr = SVDDecomposeTransMatrix(n1, n2, add_diag=True)
    tr.to_eval_mode()
    for name in ("left", "right", "left_inv", "right_inv"):
        getattr(tr, f"matrix_{name}").data = T(g[name]).clone()
    tr.use_diag = False
    fl = FlatQuantizedLinear(_args(), lin)
    with torch.no_grad():
        fl.act_quantizer.clip_factor_a_max.fill_(float(g["ca_max"]))
        fl.act_quantizer.clip_factor_a_min.fill_(float(g["ca_min"]))
        fl.clip_factor_w_max.copy_(T(g["cw_max"]))
        fl.clip_factor_w_min.copy_(T(g["cw_min"]))
        fl.reparameterize(qa_trans=tr)
        assert torch.allclose(lin.weight.data, T(g["w_rep"]), atol=1e-6)
        wq = WeightQuantizer()
        wq.configure(4, perchannel=True, sym=True, mse=False)
        wq.find_params(lin.weight.data)
        lin.weight.data = wq.quantize(lin.weight.data)
        assert torch.equal(wq.scale.reshape(-1, 1), T(g["s_w"]))
        fl, tr = fl.to(dev).half(), tr.to(dev)
        x = T(g["x"]).half().to(dev)
        y = fl(tr(x))  # transform kernel, then quant+GEMM+dequant kernel chain
        y_f = fl.forward_fused(x, tr)  # one library call
    y_ref = g["y"]
    for out in (y, y_f):
        err = np.abs(out.float().cpu().numpy() - y_ref)
        assert err.max() <= 0.03 * np.abs(y_ref).max() and err.mean() <= 0.005 * np.abs(y_ref).max()
    with pytest.raises(RuntimeError):
        fl(T(g["x"]).half())  # CPU activations: no fallback


@pytest.mark.parametrize("case", ["flatlinear_256_svd", "flatlinear_224_inv_bias", "flatlinear_4096_svd"])
def test_offline_weight_preparation_on_the_gpu(ops, dev, golden, case):
    """reparameterize (fp64 Kronecker on the weight + LWC clip) -> RTN -> 4-bit pack, with the module and
    its weight RESIDENT ON THE GPU (the path a 70B / V3-scale model takes), against the reference-generated fixture:
    the re-parameterised weight to 1e-6, the RTN scales bit for bit, the codes and the packed bytes exactly."""
    from flatquant_b200.flat_linear import FlatQuantizedLinear
    from flatquant_b200.quant_utils import WeightQuantizer
    g = golden(case)
    n1, n2 = g["left"].shape[0], g["right"].shape[0]
    n = g["w0"].shape[0]
    has_bias = g["bias"].size > 0
    lin = torch.nn.Linear(n1 * n2, n, bias=has_bias)
    lin.weight.data = T(g["w0"]).clone()
    if has_bias:
        lin.bias.data = T(g["bias"]).clone()
    fl = FlatQuantizedLinear(_args(), lin).to(dev)
    with torch.no_grad():
        fl.clip_factor_w_max.copy_(T(g["cw_max"]).to(dev))
        fl.clip_factor_w_min.copy_(T(g["cw_min"]).to(dev))
        fl.reparameterize(qa_trans=[T(g["left_inv"]).to(dev), T(g["right_inv"]).to(dev)])
        assert fl.linear.weight.is_cuda
        assert torch.allclose(fl.linear.weight.data.cpu(), T(g["w_rep"]), atol=1e-6)
        wq = WeightQuantizer()
        wq.configure(4, perchannel=True, sym=True, mse=False)
        wq.find_params(fl.linear.weight.data)
        fl.linear.weight.data = wq.quantize(fl.linear.weight.data)
        # the fp64 products are summed in another order on the GPU: the fp32 weight (and hence a scale, or a code
        # sitting exactly on a rounding tie) may differ in the last bit at K = 4096; the small fixtures are bit-equal
        big = n1 * n2 >= 4096
        if big:
            assert torch.allclose(wq.scale.reshape(-1, 1).cpu(), T(g["s_w"]), rtol=1e-6, atol=0)
        else:
            assert torch.equal(wq.scale.reshape(-1, 1).cpu(), T(g["s_w"]))
        packed, s_w = fl.pack_weights(weight_scales=wq.scale)        # flatq_pack_i4 on the device
    assert packed.is_cuda
    codes = O.unpack_i4(packed.cpu().numpy())
    diff = np.abs(codes.astype(np.int32) - g["w_codes"].astype(np.int32))
    assert diff.max() <= (1 if big else 0) and (diff != 0).mean() <= (1e-5 if big else 0.0)
    if not big:
        assert np.array_equal(packed.cpu().numpy(), g["w_packed"])


def test_operator_api_bf16_activations_and_no_silent_library_fallback(ops, dev):
    """the operator-level transforms (flat_utils.kronecker_matmul, *TransMatrix.forward) take bf16
    inference activations on the sm_100a kernel (exact bf16 -> fp16 conversion on the way in, result rounded to bf16),
    and refuse 16-bit activations whose factor sizes have no kernel instead of quietly calling cuBLAS."""
    from flatquant_b200 import flat_utils as FU
    from flatquant_b200.trans_utils import _matmul_last
    torch.manual_seed(11)
    n1, n2 = 64, 64
    x = torch.randn(3, 7, n1 * n2).to(torch.bfloat16)
    L = torch.randn(n1, n1) / n1 ** 0.5
    R = torch.randn(n2, n2) / n2 ** 0.5
    calls = []
    real = ops.kron_transform
    try:
        ops.kron_transform = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        with torch.no_grad():
            z = FU.kronecker_matmul(x.to(dev), L.to(dev), R.to(dev))
            w = _matmul_last(x.to(dev).reshape(3, 7, 32, 128), R.new_tensor(torch.randn(128, 128) / 128 ** 0.5).to(dev))
    finally:
        ops.kron_transform = real
    assert len(calls) == 2 and z.dtype == torch.bfloat16 and z.shape == x.shape and w.dtype == torch.bfloat16
    ref = O.kronecker_matmul(x.float(), L.to(torch.bfloat16).float(), R.to(torch.bfloat16).float())
    err = (z.float().cpu() - ref).abs().max() / ref.abs().max()
    assert err < 2.0 ** -7, err   # one bf16 rounding of the output (2^-9 relative) + the fp16 intermediate
    with torch.no_grad(), pytest.raises(RuntimeError, match="no sm_100a kernel"):
        FU.kronecker_matmul(torch.randn(2, 14 * 15, device=dev).half(), torch.eye(14, device=dev), torch.eye(15, device=dev))
    # fp32 activations are the reference's own torch path (calibration precision), untouched
    with torch.no_grad():
        z32 = FU.kronecker_matmul(x.float().to(dev), L.to(dev), R.to(dev))
    assert z32.dtype == torch.float32


def test_deploy_api_chain(dev):
    """benchmarks/qlinear_benchmark.py:70-74: Sequential(OnlineTrans('matmul'), Quantizer, Linear4bit)."""
    from flatquant_b200 import deploy
    from flatquant_b200.deploy.nn import Linear4bit, OnlineTrans, Quantizer
    torch.manual_seed(0)
    k, n, bsz, seq = 4096, 512, 2, 24
    base = torch.nn.Linear(k, n, bias=False).half()
    base.weight.data = torch.randint(-8, 7, (n, k)).half()
    s_w = torch.ones((n, 1), dtype=torch.float16)
    mod = torch.nn.Sequential(OnlineTrans(k, trans="matmul"), Quantizer(input_clip_ratio=1.0),
                              Linear4bit.from_float(base, weight_scales=s_w)).to(dev)
    ot = mod[0]
    ot.left_matrix.data = (torch.randn(64, 64) / 8).half().to(dev)
    ot.right_matrix.data = (torch.randn(64, 64) / 8).half().to(dev)
    x = torch.rand(bsz, seq, k).half()
    y = mod(x.to(dev))
    assert y.shape == (bsz, seq, n) and y.dtype == torch.float16   # leading dims kept, as deploy.matmul / sym_dequant do
    y = y.reshape(bsz * seq, n)
    packed = ot(x.to(dev))
    # deploy/functional/online_trans.py:121-122: quantized_x [bsz, seq, K/2], scales_x [bsz, 1, seq]
    assert isinstance(packed, deploy.PackedQuantizedTensor) and packed.quantized_x.shape == (bsz, seq, k // 2)
    assert packed.scales_x.shape == (bsz, 1, seq)
    codes = O.unpack_i4(packed.quantized_x.reshape(bsz * seq, -1).cpu().numpy())
    acc = O.gemm_i4(codes, base.weight.data.float().numpy())
    ref = O.sym_dequant(acc, packed.scales_x.float().cpu().numpy().reshape(-1), np.ones(n, np.float32)).astype(np.float16)
    assert np.array_equal(y.cpu().numpy(), ref)
    # int-only chain of the reference API
    acc_gpu = deploy.matmul(packed.quantized_x, mod[2].weight)
    assert acc_gpu.shape == (bsz, seq, n) and np.array_equal(acc_gpu.reshape(-1, n).cpu().numpy(), acc)
    y2 = deploy.sym_dequant(acc_gpu, packed.scales_x, mod[2].weight_scales)
    assert np.array_equal(y2.reshape(-1, n).cpu().numpy(), ref)


def test_online_trans_single_matrix_is_the_reference_block_path(dev):
    """OnlineTrans(decompose=False) -- the o_proj transform over the heads axis -- takes the reference's 4-D
    [bsz, seq, head_dim, heads] input (modeling_llama.py:146: o_proj_trans(attn.transpose(-1, -2))): ONE scale per token
    over the whole block, codes of (b_t @ inv) packed in the transposed (heads-major) order, quantized_x reshaped to
    (bsz, seq, -1, heads), scales_x to (bsz, 1, seq)  (deploy/functional/online_trans.py:124-139, block_matmul.py:29-104)."""
    from flatquant_b200 import deploy
    from flatquant_b200.deploy.nn import OnlineTrans
    torch.manual_seed(5)
    bsz, seq, heads, hd = 2, 9, 32, 128
    ot = OnlineTrans(heads, trans="matmul", decompose=False).to(dev)
    inv = (torch.eye(heads) * 0.8 + torch.randn(heads, heads) / 16).half()
    ot.right_matrix.data = inv.to(dev)
    attn = torch.randn(bsz, seq, heads, hd).half()
    packed = ot(attn.transpose(-1, -2).contiguous().to(dev))
    assert isinstance(packed, deploy.PackedQuantizedTensor)
    assert packed.quantized_x.shape == (bsz, seq, hd * heads // 2 // heads, heads) and packed.scales_x.shape == (bsz, 1, seq)
    # oracle: z_t = b_t @ inv with b_t [head_dim, heads]; per-token scale over the whole block; codes of z_t^T row-major
    z = attn.transpose(-1, -2).float() @ inv.float()                     # [bsz, seq, hd, heads]
    zt = z.transpose(-1, -2).reshape(bsz * seq, heads * hd)               # heads-major
    s = O.act_scale(zt, 1.0, 1.0)                                          # sigmoid(1.0) clip factors (kron_matmul.py:86-92)
    codes = O.sym_codes(zt, s).numpy()
    got = O.unpack_i4(packed.quantized_x.reshape(bsz * seq, -1).cpu().numpy())
    flips = (got != codes)
    assert flips.mean() <= 0.006 and np.abs(got.astype(np.int32) - codes).max() <= 1
    assert torch.allclose(packed.scales_x.reshape(-1).float().cpu(), s.reshape(-1), rtol=2e-3)


def test_error_behaviour(ops, dev):
    a = torch.zeros((4, 24), dtype=torch.uint8, device=dev)  # K = 48: not a multiple of 64
    with pytest.raises(RuntimeError):
        ops.gemm_i4(a, a)
    with pytest.raises(RuntimeError):
        ops.pack_i4(torch.zeros((2, 16), dtype=torch.int8))  # CPU tensor
    with pytest.raises(RuntimeError):
        ops.kron_transform(torch.zeros((2, 15 * 24), dtype=torch.float16, device=dev),
                           torch.zeros((15, 15), dtype=torch.float16, device=dev),
                           torch.zeros((24, 24), dtype=torch.float16, device=dev), 15, 24)  # n2 % 16 != 0


def test_gemm_i4_repeatable(ops, dev):
    """Race detector: the same large GEMM 12 times must give the identical (exact) result every time."""
    rng = np.random.default_rng(99)
    m, n, k = 1024, 2048, 4096
    a = rng.integers(-8, 8, (m, k)).astype(np.int8)
    b = rng.integers(-8, 8, (n, k)).astype(np.int8)
    ap, bp = T(O.pack_i4(a)).to(dev), T(O.pack_i4(b)).to(dev)
    ref = torch.from_numpy(O.gemm_i4(a, b)).to(dev)
    for _ in range(12):
        assert torch.equal(ops.gemm_i4(ap, bp), ref)


# ------------------------------------------------------------------ a14: qattention side path
@pytest.mark.parametrize("n", [32, 128, 1024, 4096])
def test_hadamard_golden(ops, dev, golden, n):
    g = golden("hadamard")
    x = T(g[f"x_{n}"]).half()
    y = ops.hadamard(x.to(dev), 1.0 / n ** 0.5).float().cpu().numpy()
    ref = O.hadamard_transform(x.float().numpy(), 1.0 / n ** 0.5)
    # FHT tolerance rule of third-party/fast-hadamard-transform/tests: a few fp16 ulps of the output magnitude
    assert np.abs(y - ref).max() <= 2 ** -9 * max(1.0, np.abs(ref).max())
    assert np.abs(y - g[f"y_{n}"]).max() <= 2e-2  # golden is from the fp32 input; x was rounded to fp16 here


@pytest.mark.parametrize("n", [12 * 8, 20 * 16, 28 * 32, 40 * 8, 36 * 8, 52 * 8, 60 * 8, 108 * 8, 140 * 8, 156 * 8, 172 * 8,
                               172 * 64, 28 * 512])
def test_hadamard_kn_golden(ops, dev, golden, n):
    """matmul_hadU_cuda for n = K * 2^m on the fused kernel (FHT segments + the reference's K x K table) against the
    reference's own matmul_hadU output (tests/golden/make_golden_hadamard_kn.py) and the oracle's restatement of the
    deployed form (fp16 intermediate, deploy/functional/online_trans.py:144-151)."""
    from flatquant_b200.deploy.functional.hadamard import get_hadK, matmul_hadU_cuda
    g = golden("hadamard_kn")
    x = T(g[f"x_{n}"])
    hk, k = get_hadK(n)
    y = matmul_hadU_cuda(x.to(dev), hk.to(dev), k).float().cpu().numpy()
    ref = O.matmul_hadU_kn(x.float().numpy(), hk.numpy(), 1.0 / n ** 0.5)
    # vs the same rounding points: one fp16 rounding of the output (+ fp32 summation order)
    assert np.abs(y - ref).max() <= 2 ** -10 * max(1.0, np.abs(ref).max()) * 1.5
    # vs the reference's fp32 result: the fp16 intermediate contributes sqrt(K) half-ulps
    assert np.abs(y - g[f"y_{n}"]).max() <= 2 ** -9 * k ** 0.5 * max(1.0, np.abs(ref).max())


def test_hadamard_kn_full_width_properties(ops, dev):
    """Llama-3-8B down_proj in the Hadamard baseline mode: 2048 x 14336 = 28 x 512.  H.H^T = n.I: applying the transform
    and then its transpose (hadK^T, same FHT) is the identity; Parseval per row."""
    from flatquant_b200.deploy.functional.hadamard import get_hadK, matmul_hadU_cuda
    torch.manual_seed(3)
    n = 14336
    x = torch.randn(2048, n, device=dev).half()
    hk, k = get_hadK(n)
    hkt, _ = get_hadK(n, transpose=True)
    y = matmul_hadU_cuda(x, hk.to(dev), k)
    z = matmul_hadU_cuda(y, hkt.to(dev), k)
    assert (z.float() - x.float()).abs().max() < 3e-2
    assert torch.allclose(y.float().pow(2).sum(1), x.float().pow(2).sum(1), rtol=1e-2)


def test_hadamard_properties_large(ops, dev):
    """n = 32768 (the FHT maximum): H.H = n.I (round trip) and Parseval, on 64 rows."""
    torch.manual_seed(0)
    x = torch.randn(64, 32768, device=dev).half()
    y = ops.hadamard(x, 1.0 / 32768 ** 0.5)
    z = ops.hadamard(y, 1.0 / 32768 ** 0.5)
    assert (z.float() - x.float()).abs().max() < 2e-2
    assert torch.allclose(y.float().pow(2).sum(1), x.float().pow(2).sum(1), rtol=1e-2)


def test_kv_quant_golden_bit_exact(ops, dev, golden):
    g = golden("kv_asym")
    x = T(g["x"]).reshape(-1, g["x"].shape[-2], 128)  # [tokens, heads, 128]
    q, par = ops.kv_quant(x.to(dev))
    assert np.array_equal(q.cpu().numpy(), g["packed"].reshape(q.shape))
    assert np.array_equal(par[..., 0].cpu().numpy(), g["scale"].reshape(par.shape[:2]))
    assert np.array_equal(par[..., 1].cpu().numpy(), g["zero"].reshape(par.shape[:2]))
    torch.manual_seed(1)
    x2 = (torch.randn(300, 8, 128) * 2).half()
    q2, par2 = ops.kv_quant(x2.to(dev))
    pk, s, z = O.kv_asym_quant(x2)
    assert torch.equal(q2.cpu(), pk) and torch.equal(par2[..., :1].cpu(), s) and torch.equal(par2[..., 1:].cpu(), z)


def test_paged_kv_append_and_decode(ops, dev):
    """init (prefill) + append (decode) into the paged INT4 cache, then single-query attention, vs the oracle."""
    torch.manual_seed(2)
    B, Hq, Hkv, D, PS, L, layer = 3, 8, 2, 128, 16, 2, 1
    lens = [37, 16, 5]
    pages_per = [(n + PS - 1) // PS for n in lens]
    n_pages = sum(pages_per) + 2
    kv_data = torch.zeros(n_pages, L, 2, Hkv, PS, D // 2, dtype=torch.uint8, device=dev)
    kv_param = torch.zeros(n_pages, L, 2, Hkv, PS, 2, dtype=torch.float16, device=dev)
    perm = torch.randperm(n_pages)[:sum(pages_per)].tolist()  # scattered page ids
    indptr, indices, slots, last = [0], [], [], []
    ks, vs = [], []
    for b, n in enumerate(lens):
        pg = perm[indptr[-1]:indptr[-1] + pages_per[b]]
        indices += pg
        indptr.append(indptr[-1] + pages_per[b])
        last.append(n - (pages_per[b] - 1) * PS)
        slots += [pg[t // PS] * PS + t % PS for t in range(n)]
        ks.append((torch.randn(n, Hkv, D) * 1.5).half())
        vs.append(torch.randn(n, Hkv, D).half())
    k_all, v_all = torch.cat(ks).to(dev), torch.cat(vs).to(dev)
    slot_t = torch.tensor(slots, dtype=torch.int32, device=dev)
    # all but the last token of every sequence in one "init" call, the last one as a decode-time append
    is_last = torch.zeros(len(slots), dtype=torch.bool)
    pos = 0
    for n in lens:
        is_last[pos + n - 1] = True
        pos += n
    for sel, src in ((0, k_all), (1, v_all)):
        ops.kv_quant(src[~is_last.to(dev)].contiguous(), kv_data, kv_param, slot_t[~is_last.to(dev)].contiguous(), layer, sel)
        ops.kv_quant(src[is_last.to(dev)].contiguous(), kv_data, kv_param, slot_t[is_last.to(dev)].contiguous(), layer, sel)
    q = torch.randn(B, Hq, D).half()
    o = ops.kv_decode_i4(q.to(dev), kv_data, kv_param, torch.tensor(indptr, dtype=torch.int32, device=dev),
                         torch.tensor(indices, dtype=torch.int32, device=dev), torch.tensor(last, dtype=torch.int32, device=dev),
                         layer=layer).float().cpu()
    assert not kv_data[:, 0].any()  # the other layer is untouched
    for b, n in enumerate(lens):
        kp, ksc, kz = O.kv_asym_quant(ks[b])
        vp, vsc, vz = O.kv_asym_quant(vs[b])
        rep = Hq // Hkv
        ref = O.decode_attention_i4(q[b:b + 1], kp.repeat_interleave(rep, 1)[None], ksc.repeat_interleave(rep, 1)[None],
                                    kz.repeat_interleave(rep, 1)[None], vp.repeat_interleave(rep, 1)[None],
                                    vsc.repeat_interleave(rep, 1)[None], vz.repeat_interleave(rep, 1)[None])[0]
        assert (o[b] - ref).abs().max() < 5e-3  # fp16 output + fp16-vs-fp32 dequant arithmetic


def test_kv_cache_module_prefill_then_decode(dev):
    """deploy.transformers.MultiLayerPagedKVCache4Bit protocol (kv_cache.py:231-359): prefill update returns the
    transformed fp16 states, decode update returns a closure running attention over the INT4 pages."""
    from flatquant_b200.deploy.transformers import MultiLayerPagedKVCache4Bit
    torch.manual_seed(5)
    B, T, H, D, L = 2, 21, 4, 128, 2
    cache = MultiLayerPagedKVCache4Bit(B, 16, 64, dev, L, H, D, trans="had")
    k = torch.randn(B, T, H, D, device=dev).half()
    v = torch.randn(B, T, H, D, device=dev).half()
    for layer in range(L):
        kt, vt = cache.update(k, v, layer)
        assert kt.shape == k.shape and torch.equal(vt, v)
    assert cache.seen_tokens == T and cache.get_seq_length() == T
    k1 = torch.randn(B, 1, H, D, device=dev).half()
    v1 = torch.randn(B, 1, H, D, device=dev).half()
    q = torch.randn(B, 1, 2 * H, D, device=dev).half()  # GQA: 2 query heads per kv head
    for layer in range(L):
        attn = cache.update(k1, v1, layer)
        o = attn(q)
        assert o.shape == (B, 1, 2 * H, D)
        if layer == L - 1:
            had = lambda t: torch.from_numpy(O.hadamard_transform(t.float().cpu().numpy(), 1 / D ** 0.5)).half()  # noqa: E731
            kk = torch.cat([had(k), had(k1)], 1)
            vv = torch.cat([v.cpu(), v1.cpu()], 1)
            kp, ks, kz = O.kv_asym_quant(kk)
            vp, vs_, vz = O.kv_asym_quant(vv)
            r = lambda t: t.repeat_interleave(2, 2)  # noqa: E731
            ref = O.decode_attention_i4(had(q)[:, 0], r(kp), r(ks), r(kz), r(vp), r(vs_), r(vz))
            assert (o[:, 0].float().cpu() - ref).abs().max() < 2e-2


def test_kernel_level_api(dev):
    """deploy.kernels.kron_matmul / block_matmul signatures (kron_matmul.py:192, block_matmul.py:231)."""
    from flatquant_b200.deploy.kernels.block_matmul import block_matmul
    from flatquant_b200.deploy.kernels.kron_matmul import kron_matmul
    torch.manual_seed(6)
    B, n1, n2 = 10, 64, 64
    L = (torch.randn(n1, n1) / 8).half()
    R = (torch.randn(n2, n2) / 8).half()
    x = torch.randn(B, n1, n2).half()
    pk = kron_matmul(L.t().contiguous().to(dev), x.to(dev), R.to(dev), 5, 4.0, 4.0)
    q_em, _ = _emulated_codes(x.reshape(B, -1), L, R, float(torch.sigmoid(torch.tensor(4.0))), float(torch.sigmoid(torch.tensor(4.0))), "xr")
    d = np.abs(O.unpack_i4(pk.quantized_x.cpu().numpy()) - q_em)
    assert pk.scales_x.shape == (B, 1) and d.max() <= 1 and (d != 0).mean() < 2e-3
    heads, hd = 32, 128
    a = torch.randn(B, hd, heads).half()  # [B, M=head_dim, N=heads]
    o_t = (torch.randn(heads, heads) / heads ** 0.5).half()
    pb = block_matmul(a.to(dev), o_t.to(dev), 5, 4.0, 4.0)
    z = torch.matmul(a.float(), o_t.float()).transpose(1, 2).reshape(B, -1)  # transposed packing order
    c = float(torch.sigmoid(torch.tensor(4.0)))
    s = torch.maximum((z.amin(1).clamp(max=0) * c).abs(), z.amax(1).clamp(min=0) * c) / 7
    qz = torch.clamp(torch.round(z / s[:, None]), -8, 7).numpy().astype(np.int32)
    d = np.abs(O.unpack_i4(pb.quantized_x.cpu().numpy()) - qz)
    assert d.max() <= 1 and (d != 0).mean() < 0.02


@pytest.mark.parametrize("m,k,n,n1,n2", [(1, 4096, 4096, 64, 64), (16, 4096, 6144, 64, 64), (64, 14336, 4096, 112, 128),
                                         (5, 256, 128, 16, 16), (2, 4096, 1000, 64, 64), (2, 14336, 4096, 112, 128), (9, 14336, 4096, 112, 128),
                                         (17, 4096, 4096, 64, 64), (40, 4096, 28672, 64, 64)])
def test_decode_split_k_path(ops, dev, m, k, n, n1, n2):
    """Decode shapes: M <= 2 -> dp4a GEMV kernel, M <= 32 -> weight-streaming mma.sync kernel (gemm_w4a8_skinny.cuh),
    M <= 64 -> split-K tcgen05 route (int32 atomics + finalize);
    results must be identical to the
    oracle arithmetic on the kernel's own codes, and run-to-run deterministic (integer atomics commute)."""
    torch.manual_seed(m + n)
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    wp = ops.pack_i4(wq)
    s_w = torch.rand(n, device=dev) * 0.02 + 0.01
    bias = torch.randn(n, device=dev).half()
    if m <= 2:   # linear_fused is the one-launch kernel there (test_one_launch_decode_flatlinear): pin the GEMV by itself
        fused = lambda: ops.gemm_w4a8(*ops.act_quant(x, L, R, n1, n2, 0.97, 0.99), wp, s_w, bias)   # noqa: E731
    else:
        fused = lambda: ops.linear_fused(x, L, R, n1, n2, 0.97, 0.99, wp, s_w, bias)   # noqa: E731
    y = fused()
    y_again = fused()
    assert torch.equal(y, y_again)
    xq8, sx = ops.act_quant(x, L, R, n1, n2, 0.97, 0.99)
    perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    c8 = xq8.cpu().numpy().reshape(m, k // 8, 8)
    codes = np.empty_like(c8)
    codes[:, :, perm] = c8
    acc = O.gemm_i4(codes.reshape(m, k), wq.cpu().numpy())
    ref = O.sym_dequant(acc, sx.cpu().numpy(), s_w.cpu().numpy(), bias.float().cpu().numpy()).astype(np.float16)
    assert np.array_equal(y.cpu().numpy(), ref)


@pytest.mark.parametrize("m,n,k", [(2048, 4096, 4096), (2048, 6144, 4096), (2048, 5120, 8192), (1500, 4100, 2048),
                                   (2048, 4096, 14336), (4096, 2560, 1024), (2048, 28672, 4096)])
def test_prefill_gemm_variants_bit_identical(ops, dev, m, n, k):
    """Prefill GEMM variants must agree bit for bit (integer accumulation, same fp32 epilogue): the default CTA-pair
    kernel (csrc/gemm_w4a8_2sm.cuh), the single-CTA kernel, and the single-CTA kernel with the stream-K schedule on
    the partial last wave (int32 partials between CTAs, csrc/gemm_w4a8.cuh GemmSched); all equal to the oracle
    arithmetic (sampled rows) and run-to-run deterministic."""
    import ctypes
    from flatquant_b200 import _lib
    torch.manual_seed(m + n + k)
    xq8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev)
    sx = torch.rand(m, device=dev) * 0.05 + 0.01
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    wp = ops.pack_i4(wq)
    sw = torch.rand(n, device=dev) * 0.02 + 0.01
    bias = torch.randn(n, device=dev).half()
    lib = _lib.load()
    setopt = lib.flatq_debug_set_gemm_options
    setopt.restype, setopt.argtypes = None, [ctypes.c_int, ctypes.c_int]
    try:
        y_pair = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
        y_pair2 = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
        setopt(1, 0)
        y_1sm = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
        setopt(1, 1)
        y_sk = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)   # ops passes the stream-K workspace
        y_sk2 = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
        torch.cuda.synchronize()
    finally:
        setopt(-1, -1)
    assert torch.equal(y_pair, y_pair2) and torch.equal(y_pair, y_1sm) and torch.equal(y_pair, y_sk) and torch.equal(y_sk, y_sk2)
    rows = np.unique(np.concatenate([np.arange(0, m, max(1, m // 23)), [m - 1]]))
    perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    c8 = xq8[rows].cpu().numpy().reshape(len(rows), k // 8, 8)
    codes = np.empty_like(c8)
    codes[:, :, perm] = c8
    acc = O.gemm_i4(codes.reshape(len(rows), k), wq.cpu().numpy())
    ref = O.sym_dequant(acc, sx[rows].cpu().numpy(), sw.cpu().numpy(), bias.float().cpu().numpy()).astype(np.float16)
    assert np.array_equal(y_pair[rows].cpu().numpy(), ref)


def test_moe_routed_experts_configs4(dev):
    """DeepSeek-V3 routed experts (configs[4]) at reduced width: shared w1_trans / quantiser, per-expert w1,w3,w2."""
    from flatquant_b200 import ops as _ops
    from flatquant_b200.moe import FlatQuantRoutedExperts
    torch.manual_seed(9)
    T_, dim, inter, E, top_k = 48, 256, 128, 4, 2
    (a1, b1), (a2, b2) = O.get_decompose_dim(dim), O.get_decompose_dim(inter)
    x = torch.randn(T_, dim).half()
    L1, R1 = (torch.randn(a1, a1) / a1 ** 0.5).half(), (torch.randn(b1, b1) / b1 ** 0.5).half()
    L2, R2 = (torch.randn(a2, a2) / a2 ** 0.5).half(), (torch.randn(b2, b2) / b2 ** 0.5).half()
    w13 = torch.randint(-8, 8, (E, 2 * inter, dim), dtype=torch.int8)
    w2 = torch.randint(-8, 8, (E, dim, inter), dtype=torch.int8)
    s13, s2 = torch.rand(E, 2 * inter) * 0.02 + 0.01, torch.rand(E, dim) * 0.02 + 0.01
    indices = torch.multinomial(torch.tensor([0.45, 0.3, 0.2, 0.05]).repeat(T_, 1), top_k)  # skewed load, distinct experts per token
    weights = torch.rand(T_, top_k)
    moe = FlatQuantRoutedExperts(torch.stack([T(O.pack_i4(w13[e].numpy())) for e in range(E)]).to(dev), s13.to(dev),
                                 torch.stack([T(O.pack_i4(w2[e].numpy())) for e in range(E)]).to(dev), s2.to(dev),
                                 (L1.to(dev), R1.to(dev)), (L2.to(dev), R2.to(dev)), (0.98, 0.98), (0.98, 0.98))
    y = moe(x.to(dev), weights.to(dev), indices.to(dev)).float().cpu()
    # oracle with the reference's structure (python loop over experts), fp32 arithmetic, raw clip = logit(0.98)
    raw = float(torch.logit(torch.tensor(0.98)))
    z = O.kronecker_matmul(x.float(), L1.float(), R1.float())
    ref = torch.zeros(T_, dim)
    for e in range(E):
        idx, top = torch.where(indices == e)
        if idx.numel() == 0:
            continue
        gu = torch.from_numpy(O.flat_linear_int(z[idx], None, None, raw, raw, w13[e].numpy(), s13[e].numpy())["y"])
        hh = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
        out = torch.from_numpy(O.flat_linear_int(hh, L2.float(), R2.float(), raw, raw, w2[e].numpy(), s2[e].numpy())["y"])
        ref[idx] += out * weights[idx, top, None]
    err = (y - ref).abs()
    # two chained quantised stages at tiny K (256, 128): a single +-1 code flip moves y by several % of max|y|
    rmax = float(ref.abs().max())
    assert err.max() <= 0.15 * rmax and err.mean() <= 0.02 * rmax, (float(err.max()), float(err.mean()), rmax)


def test_empty_and_maximum_shapes(ops, dev):
    """Edge cases: empty batch; the largest K of the configs (28672, factors (128, 224) -> HMMA transform kernel)."""
    k = 4096
    wp = torch.zeros((128, k // 2), dtype=torch.uint8, device=dev)
    y = ops.linear_fused(torch.zeros((0, k), dtype=torch.float16, device=dev), None, None, 1, k, 1.0, 1.0, wp,
                         torch.ones(128, device=dev))
    assert y.shape == (0, 128)
    assert ops.gemm_i4(torch.zeros((0, k // 2), dtype=torch.uint8, device=dev), wp).shape == (0, 128)
    torch.manual_seed(12)
    m, n, n1, n2 = 70, 256, 128, 224
    k = n1 * n2
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    s_w = torch.rand(n, device=dev) * 0.01 + 0.005
    y = ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, ops.pack_i4(wq), s_w)
    xq, _, s32 = ops.transform_quant(x, L, R, n1, n2, 0.98, 0.98)
    acc = O.gemm_i4(O.unpack_i4(xq.cpu().numpy()), wq.cpu().numpy())
    ref = O.sym_dequant(acc, s32.cpu().numpy(), s_w.cpu().numpy()).astype(np.float16)
    assert np.array_equal(y.cpu().numpy(), ref)
    assert np.abs(acc).max() < 2 ** 24  # the folded-shift fast path of the epilogue is exact only below this


@pytest.mark.parametrize("k", [2048, 2176, 4096, 5120, 7168, 8192, 14336])
@pytest.mark.parametrize("m", [1, 2, 3, 8, 9, 16, 17, 31, 32])
def test_decode_gemv_and_skinny_gemm_bit_exact(ops, dev, m, k):
    """One / two tokens: the GEMV (csrc/gemv_w4a8.cuh) -- the register-resident variant that requests its first weight
    rows before the PDL wait (K a multiple of 1024 among the instantiated widths) and the generic one (K = 2048, 2176);
    3..32 tokens: the weight-streaming mma.sync.m16n8k32 kernel (csrc/gemm_w4a8_skinny.cuh; 1, 2 and 4 eight-token
    blocks, K = 2176 = 17 super-steps leaves its warps uneven).  Against the integer arithmetic of the path, bit for
    bit; N ragged against the grid / the 16-row tile, with bias."""
    torch.manual_seed(m * 100003 + k)
    n = 4100 if k != 4096 else 28672 + 8
    xq8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev)
    sx = torch.rand(m, device=dev) * 0.05 + 0.01
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    wp = ops.pack_i4(wq)
    sw = torch.rand(n, device=dev) * 0.02 + 0.01
    bias = torch.randn(n, device=dev).half()
    y = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
    y2 = ops.gemm_w4a8(xq8, sx, wp, sw, bias, torch.float16)
    perm = torch.tensor([0, 2, 4, 6, 1, 3, 5, 7], device=dev)
    codes = torch.empty_like(xq8).view(m, k // 8, 8)
    codes[:, :, perm] = xq8.view(m, k // 8, 8)          # the operand is stored K-permuted
    acc = codes.view(m, k).float() @ wq.float().t()      # exact in fp32: |sum| <= 64 K < 2^24
    ref = ((acc * sx[:, None]) * sw[None, :] + bias.float()[None, :]).half()
    assert torch.equal(y, ref) and torch.equal(y, y2)


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256])
def test_hadamard_short_rows_in_warp_lanes(ops, dev, n):
    """n <= 256: a row lives in n / 8 lanes of a warp (fht_warp_kernel, no shared memory); ragged row counts (the last
    warp is partly empty), against the oracle's transform, and an involution: H(H(x)) / n == x."""
    torch.manual_seed(n)
    for rows in (1, 37, 4099):
        x = torch.randn(rows, n).half()
        y = ops.hadamard(x.to(dev), 1.0 / n ** 0.5)
        ref = O.hadamard_transform(x.float().numpy(), 1.0 / n ** 0.5)
        assert np.abs(y.float().cpu().numpy() - ref).max() <= 2 ** -9 * max(1.0, np.abs(ref).max())
        back = ops.hadamard(y, 1.0 / n ** 0.5).float().cpu()
        assert float((back - x.float()).abs().max()) <= 2 ** -8 * max(1.0, float(x.float().abs().max()))


@pytest.mark.parametrize("m,n1,n2,n,sides", [(1, 64, 64, 6144, "lr"), (2, 64, 64, 4100, "lr"), (1, 32, 128, 4096, "l"), (2, 32, 128, 4100, "l"),
                                             (1, 64, 128, 1032, "lr"),
                                             (2, 32, 64, 520, "lr"), (1, 64, 64, 4096, "r"), (2, 64, 128, 28672, "lr")])
def test_one_launch_decode_flatlinear(ops, dev, m, n1, n2, n, sides):
    """csrc/flatlinear_decode.cuh: for one or two tokens flatq_linear_fused is ONE launch -- every CTA of the weight-stream
    GEMV computes the Kronecker transform and the quantisation of the token(s) itself (mma.sync, the rounding points of the
    tcgen05 kernel: left factor first, one fp16 rounding per factor).  CTA 0 dumps its codes and scales (debug hook):
    (1) the output is EXACTLY the integer arithmetic of the path on those codes; (2) the codes equal the CPU model of the
    rounding points and the two-launch path up to rare +-1 flips (fp32 summation order inside the tensor cores);
    (3) an all-zero token gives scale 1 and zero codes."""
    import ctypes
    from flatquant_b200 import _lib
    torch.manual_seed(m * 7919 + n1 * 131 + n2 + n)
    k = n1 * n2
    x = torch.randn(m, k).half()
    if m == 2:
        x[1] = 0
    L = (torch.randn(n1, n1) / n1 ** 0.5).half() if "l" in sides else None
    R = (torch.randn(n2, n2) / n2 ** 0.5).half() if "r" in sides else None
    wq = torch.randint(-8, 8, (n, k), dtype=torch.int8, device=dev)
    wp = ops.pack_i4(wq)
    s_w = torch.rand(n, device=dev) * 0.02 + 0.01
    bias = torch.randn(n, device=dev).half()
    cmax, cmin = 0.97, 0.99
    Ld, Rd = (None if L is None else L.to(dev)), (None if R is None else R.to(dev))
    codes = torch.full((m, k), 99, dtype=torch.int8, device=dev)
    sx = torch.zeros(m, dtype=torch.float32, device=dev)
    lib = _lib.load()
    dump = lib.flatq_debug_set_decode_dump
    dump.restype, dump.argtypes = None, [ctypes.c_void_p, ctypes.c_void_p]
    try:
        dump(codes.data_ptr(), sx.data_ptr())
        y = ops.linear_fused(x.to(dev), Ld, Rd, n1, n2, cmax, cmin, wp, s_w, bias)
        torch.cuda.synchronize()
    finally:
        dump(None, None)
    assert int(codes.max()) <= 7 and int(codes.min()) >= -8, "the one-launch kernel did not run (codes not dumped)"
    y_again = ops.linear_fused(x.to(dev), Ld, Rd, n1, n2, cmax, cmin, wp, s_w, bias)
    assert torch.equal(y, y_again)
    # (1) exact given the codes: the GEMV on the dumped operand (bit-exact against the integer oracle elsewhere)
    assert torch.equal(y, ops.gemm_w4a8(codes, sx, wp, s_w, bias, torch.float16))
    # (2) codes: CPU rounding-point model (left factor first) and the two-launch path
    perm = torch.tensor([0, 2, 4, 6, 1, 3, 5, 7])
    logical = torch.empty(m, k // 8, 8, dtype=torch.int8)
    logical[:, :, perm] = codes.cpu().view(m, k // 8, 8)
    logical = logical.view(m, k).numpy().astype(np.int32)
    eye1, eye2 = torch.eye(n1).half(), torch.eye(n2).half()
    q_em, s_em = _emulated_codes(x, eye1 if L is None else L, eye2 if R is None else R, cmax, cmin, "lx")
    d = np.abs(logical - q_em)
    assert d.max() <= 1 and (d != 0).mean() < 1e-3, (d.max(), (d != 0).mean())
    assert torch.allclose(sx.cpu(), s_em, rtol=1e-3)
    xq8, sx2 = ops.act_quant(x.to(dev), Ld, Rd, n1, n2, cmax, cmin)
    d2 = (xq8.int() - codes.int()).abs()
    assert int(d2.max()) <= 1 and float((d2 != 0).float().mean()) < 1e-3
    assert torch.allclose(sx, sx2, rtol=1e-3)
    # (3) the all-zero token
    if m == 2:
        assert sx[1].item() == 1.0 and not codes[1].any() and torch.equal(y[1], bias)


@pytest.mark.parametrize("m,n1,n2,inter", [(1, 64, 64, 14336), (2, 64, 64, 1032), (1, 64, 128, 520), (2, 32, 64, 264)])
def test_one_launch_decode_gate_up_silu(ops, dev, m, n1, n2, inter):
    """flatq_linear_fused_silu at one / two tokens: the same kernel with the (gate, up) pair epilogue; exact given its codes."""
    import ctypes
    from flatquant_b200 import _lib
    torch.manual_seed(m + n1 + n2 + inter)
    k = n1 * n2
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    w_il = ops.pack_i4(torch.randint(-8, 8, (2 * inter, k), dtype=torch.int8, device=dev))
    s_il = torch.rand(2 * inter, device=dev) * 0.05 + 0.02
    codes = torch.full((m, k), 99, dtype=torch.int8, device=dev)
    sx = torch.zeros(m, dtype=torch.float32, device=dev)
    lib = _lib.load()
    dump = lib.flatq_debug_set_decode_dump
    dump.restype, dump.argtypes = None, [ctypes.c_void_p, ctypes.c_void_p]
    try:
        dump(codes.data_ptr(), sx.data_ptr())
        y = ops.linear_fused_silu(x, L, R, n1, n2, 0.98, 0.96, w_il, s_il)
        torch.cuda.synchronize()
    finally:
        dump(None, None)
    assert int(codes.max()) <= 7 and int(codes.min()) >= -8, "the one-launch kernel did not run"
    assert y.shape == (m, inter) and torch.equal(y, ops.gemm_w4a8_silu(codes, sx, w_il, s_il))
    xq8, sx2 = ops.act_quant(x, L, R, n1, n2, 0.98, 0.96)
    d2 = (xq8.int() - codes.int()).abs()
    assert int(d2.max()) <= 1 and float((d2 != 0).float().mean()) < 1e-3 and torch.allclose(sx, sx2, rtol=1e-3)
    # larger batches take the two-launch route with the same result as its parts
    x8 = torch.randn(40, k, device=dev).half()
    q8, s8 = ops.act_quant(x8, L, R, n1, n2, 0.98, 0.96)
    assert torch.equal(ops.linear_fused_silu(x8, L, R, n1, n2, 0.98, 0.96, w_il, s_il), ops.gemm_w4a8_silu(q8, s8, w_il, s_il))
