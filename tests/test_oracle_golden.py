"""Pin the CPU oracle against fixtures produced by the reference's own Python code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import flatq_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_decompose_dim(golden):
    g = golden("decompose_dim")
    for n, d in zip(g["n"], g["dims"]):
        assert O.get_decompose_dim(int(n)) == (int(d[0]), int(d[1]))
    assert O.get_decompose_dim(4096) == (64, 64)
    assert O.get_decompose_dim(14336) == (112, 128)


def test_pack_unpack_bit_exact(golden):
    g = golden("pack_i4")
    assert np.array_equal(O.pack_i4(g["q"]), g["packed"])
    assert np.array_equal(O.unpack_i4(g["packed"]), g["unpacked"])
    # element 2j -> low nibble, element 2j+1 -> high nibble, two's complement
    assert O.pack_i4(np.array([[-8, 7]]))[0, 0] == 0x78
    assert O.pack_i4(np.array([[-1, 1]]))[0, 0] == 0x1F


@pytest.mark.parametrize("dims", ["16x16", "14x16", "64x64"])
def test_kronecker(golden, dims):
    g = golden(f"kron_{dims}")
    z = O.kronecker_matmul(T(g["x"]), T(g["left"]), T(g["right"]))
    assert torch.allclose(z, T(g["z"]), rtol=1e-5, atol=1e-5)
    # definition: x @ kron(L, R)
    ref = T(g["x"]).reshape(-1, g["x"].shape[-1]).double() @ torch.kron(T(g["left"]).double(), T(g["right"]).double())
    assert torch.allclose(z.reshape(ref.shape).double(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name,kw", [("lac", dict(clip_max=4.0, clip_min=2.0)), ("plain", {}),
                                     ("ratio", dict(clip_ratio=0.9))])
def test_act_quant(golden, name, kw):
    g = golden("act_quant")
    x = T(g["x"])
    assert torch.equal(O.act_scale(x, **kw), T(g[f"{name}_scale"]))
    assert torch.equal(O.act_fake_quant(x, **kw), T(g[f"{name}_fq"]))
    xh = x.half()
    assert torch.equal(O.act_scale(xh, **kw), T(g[f"{name}_scale_h"]))
    assert torch.equal(O.act_fake_quant(xh, **kw), T(g[f"{name}_fq_h"]))


def test_weight_rtn(golden):
    g = golden("weight_rtn")
    q, s = O.weight_rtn(T(g["w"]))
    assert torch.equal(s, T(g["scale"]))
    assert torch.equal(q.float() * s, T(g["w_fq"]))


def test_sym_quant_half(golden):
    g = golden("sym_quant_half")
    assert np.array_equal(O.sym_quant_half(g["x"], g["scale"]).astype(np.float16), g["q"])


@pytest.mark.parametrize("case", ["flatlinear_256_svd", "flatlinear_224_inv_bias", "flatlinear_4096_svd"])
def test_flat_linear_path(golden, case):
    g = golden(case)
    x = T(g["x"])
    L, R = T(g["left"]), T(g["right"])
    # offline: reparameterize + RTN + pack
    w_rep = O.reparameterize_weight(T(g["w0"]), T(g["left_inv"]), T(g["right_inv"]), T(g["cw_max"]), T(g["cw_min"]))
    assert torch.allclose(w_rep, T(g["w_rep"]), rtol=0, atol=1e-6)
    codes, s_w = O.weight_rtn(T(g["w_rep"]))
    assert np.array_equal(codes.numpy(), g["w_codes"])
    assert torch.equal(s_w, T(g["s_w"]))
    assert np.array_equal(O.pack_i4(codes.numpy()), g["w_packed"])
    # inverse factors are (M^-1)^T
    assert torch.allclose(O.inverse_transpose(L), T(g["left_inv"]), atol=2e-5)
    # online
    bias = g["bias"] if g["bias"].size else None
    r = O.flat_linear_int(x, L, R, float(g["ca_max"]), float(g["ca_min"]), g["w_codes"], g["s_w"], bias)
    assert torch.allclose(r["z"], T(g["z"]).reshape(r["z"].shape), atol=1e-5)
    zq = (r["s_x"] * torch.from_numpy(r["xq"]).float())
    assert torch.allclose(zq, T(g["zq"]).reshape(zq.shape), atol=1e-6)
    y_ref = g["y"].reshape(r["y"].shape)
    assert np.abs(r["y"] - y_ref).max() <= 2e-5 * max(1.0, np.abs(y_ref).max())
    # fake-quant restatement too
    w_fq = T(g["w_codes"]).float() * T(g["s_w"])
    yf = O.flat_linear_fake(x, L, R, float(g["ca_max"]), float(g["ca_min"]), w_fq, T(bias) if bias is not None else None)
    assert torch.allclose(yf.reshape(y_ref.shape), T(y_ref), atol=2e-5)


def test_oproj_variant(golden):
    g = golden("flatlinear_oproj")
    w_rep = O.reparameterize_weight(T(g["w0"]), T(g["o_inv_t"]), T(g["v_inv_t"]), T(g["cw_max"]), T(g["cw_min"]))
    assert torch.allclose(w_rep, T(g["w_rep"]), atol=1e-6)
    attn = T(g["attn"])
    heads = g["o"].shape[0]
    a = attn.reshape(-1, heads, attn.shape[-1] // heads)
    z = torch.matmul(T(g["o"]).T, a).reshape(attn.shape)
    assert torch.allclose(z, T(g["z"]), atol=1e-5)
    codes, s_w = O.weight_rtn(w_rep)
    r = O.flat_linear_int(z, None, None, 4.0, 4.0, codes.numpy(), s_w.numpy())
    assert np.abs(r["y"] - g["y"].reshape(r["y"].shape)).max() < 2e-5


def test_kv_asym(golden):
    g = golden("kv_asym")
    p, s, z = O.kv_asym_quant(T(g["x"]))
    assert torch.equal(p, T(g["packed"]))
    assert torch.equal(s, T(g["scale"])) and torch.equal(z, T(g["zero"]))
    assert torch.equal(O.kv_asym_dequant(p, s, z), T(g["deq"]))


def test_hadamard(golden):
    g = golden("hadamard")
    for n in (32, 128, 1024, 4096):
        y = O.hadamard_transform(g[f"x_{n}"], scale=1.0 / np.sqrt(n))
        assert np.allclose(y, g[f"y_{n}"], rtol=1e-4, atol=1e-4)
    h = O.hadamard_matrix(8)
    assert np.array_equal(h @ h.T, 8 * np.eye(8))


def test_hadamard_kn_reference_tables(golden):
    """n = K * 2^m: the oracle's matmul_hadU_kn with the SHIPPED tables (flatquant_b200/.../data/hadamard_tables.npz)
    equals the reference's matmul_hadU (its own hard-coded tables) on every order the reference lists."""
    from flatquant_b200.deploy.functional.hadamard import get_hadK
    g = golden("hadamard_kn")
    for key in g.files:
        if not key.startswith("x_"):
            continue
        n = int(key[2:])
        hk, k = get_hadK(n)
        y = O.matmul_hadU_kn(g[key].astype(np.float32), hk.numpy(), 1.0 / np.sqrt(n), fp16_intermediate=False)
        assert np.allclose(y, g[f"y_{n}"], rtol=1e-4, atol=1e-4), (n, k)
        y16 = O.matmul_hadU_kn(g[key].astype(np.float32), hk.numpy(), 1.0 / np.sqrt(n))
        assert np.abs(y16 - g[f"y_{n}"]).max() <= 2 ** -9 * np.sqrt(k) * max(1.0, np.abs(y).max()), (n, k)


def test_int_gemm_and_dequant_models():
    rng = np.random.default_rng(0)
    xq = rng.integers(-8, 8, (5, 64)).astype(np.int8)
    wq = rng.integers(-8, 8, (7, 64)).astype(np.int8)
    acc = O.gemm_i4(xq, wq)
    assert np.array_equal(acc, xq.astype(np.int64) @ wq.astype(np.int64).T)
    sr = rng.random(5).astype(np.float16)
    sc = rng.random(7).astype(np.float16)
    y = O.sym_dequant(acc, sr, sc)
    yk = O.sym_dequant_ref_kernel(acc, sr, sc)
    assert y.shape == yk.shape == (5, 7)
    # the reference kernel's /10 truncation loses at most 10 accumulator units
    assert np.abs(yk.astype(np.float32) - y).max() <= (10 + 1) * float(sr.max()) * float(sc.max()) + 0.05 * np.abs(y).max()
