"""Generate golden fixtures by running the *reference's own Python code* on CPU.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports, unmodified, from /root/reference:
  flatquant/{flat_linear,quant_utils,trans_utils,flat_utils,function_utils}.py   (importable as-is)
  deploy/functional/quantization.py      (loaded by path: deploy/__init__ needs the unbuilt _CUDA)
  deploy/transformers/kv_cache.py        (two pure-torch functions extracted with `ast`; the module
                                          top imports the unbuilt _CUDA / fast_hadamard_transform)
  flatquant/hadamard_utils.py            (imported with a stub `fast_hadamard_transform` module;
                                          only the pure-torch `matmul_hadU` is called)
The reference has no golden vectors of its own; these files are the pin.
Outputs: tests/golden/*.npz (small, committed).
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from flatquant.flat_linear import FlatQuantizedLinear  # noqa: E402
from flatquant.flat_utils import kronecker_matmul  # noqa: E402
from flatquant.function_utils import get_decompose_dim  # noqa: E402
from flatquant.quant_utils import ActivationQuantizer, WeightQuantizer  # noqa: E402
from flatquant.trans_utils import (InvDecomposeTransMatrix, SVDDecomposeTransMatrix,  # noqa: E402
                                   SVDSingleTransMatrix)


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


dq = load_by_path("ref_deploy_quantization", f"{REF}/deploy/functional/quantization.py")


def extract_functions(path, names, namespace):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []  # drop @torch.jit.script
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
    return namespace


def args_ns(**kw):
    base = dict(w_bits=4, a_bits=4, w_asym=False, a_asym=False, lac=True, lwc=True, a_groupsize=-1)
    base.update(kw)
    return types.SimpleNamespace(**base)


def rtn(linear):
    """The 8 lines of gptq_utils.py:296-304 applied to one layer (gptq_utils itself imports tqdm/
    utils that pull in accelerate)."""
    q = WeightQuantizer()
    q.configure(4, perchannel=True, sym=True, mse=False)
    w = linear.weight.data
    q.find_params(w)
    linear.weight.data = q.quantize(w).to(w.dtype)
    return q


def flat_linear_case(tag, m, k, n, seed, trans_kind="svd", bias=False, ca=(4.0, 4.0), dtype=torch.float32):
    torch.manual_seed(seed)
    np.random.seed(seed)
    n1, n2 = get_decompose_dim(k)
    x = torch.randn(1, m, k, dtype=dtype)
    lin = torch.nn.Linear(k, n, bias=bias)
    w0 = lin.weight.data.clone()
    b0 = lin.bias.data.clone() if bias else None
    if trans_kind == "svd":
        tr = SVDDecomposeTransMatrix(n1, n2, add_diag=True)
        # make the learned factors non-trivial but well conditioned
        with torch.no_grad():
            tr.linear_diag_left.mul_(torch.linspace(0.8, 1.25, n1))
            tr.linear_diag_right.mul_(torch.linspace(1.2, 0.85, n2))
    else:
        tr = InvDecomposeTransMatrix(n1, n2, add_diag=True)
        with torch.no_grad():
            tr.linear_left.weight.add_(0.05 * torch.randn(n1, n1))
            tr.linear_right.weight.add_(0.05 * torch.randn(n2, n2))
    tr.use_diag = False  # after reparameterize_ln the diag lives in the RMSNorm (flat_utils.py:27)
    fl = FlatQuantizedLinear(args_ns(), lin)
    with torch.no_grad():
        fl.act_quantizer.clip_factor_a_max.fill_(ca[0])
        fl.act_quantizer.clip_factor_a_min.fill_(ca[1])
        fl.clip_factor_w_max.copy_(3.0 + torch.rand(n, 1))
        fl.clip_factor_w_min.copy_(3.0 + torch.rand(n, 1))
    tr.to_eval_mode()
    with torch.no_grad():
        fl.reparameterize(qa_trans=tr)
        w_rep = lin.weight.data.clone()
        wq_mod = rtn(lin)
        w_fq = lin.weight.data.clone()
        z = tr(x)
        zq = fl.act_quantizer(z)
        y = fl(z)
    s_w = wq_mod.scale.reshape(-1, 1).float()
    codes = torch.round(w_fq / s_w).to(torch.int8)
    assert torch.equal(codes.float() * s_w, w_fq)
    np.savez_compressed(
        f"{OUT}/{tag}.npz",
        x=x.numpy(), w0=w0.numpy(), bias=(b0.numpy() if bias else np.zeros(0, np.float32)),
        left=tr.matrix_left.data.numpy(), right=tr.matrix_right.data.numpy(),
        left_inv=tr.matrix_left_inv.data.numpy(), right_inv=tr.matrix_right_inv.data.numpy(),
        ca_max=np.float32(ca[0]), ca_min=np.float32(ca[1]),
        cw_max=fl.clip_factor_w_max.data.numpy(), cw_min=fl.clip_factor_w_min.data.numpy(),
        w_rep=w_rep.numpy(), w_codes=codes.numpy(), s_w=s_w.numpy(),
        w_packed=dq.pack_i4(codes).numpy(),
        z=z.numpy(), zq=zq.numpy(), y=y.numpy(),
    )
    print(tag, "y absmax", float(y.abs().max()))


def main():
    # ---- a1: get_decompose_dim table
    ns = [64, 128, 224, 256, 1024, 1536, 2048, 3584, 4096, 5120, 7168, 8192, 11008, 13824, 14336,
          16384, 18432, 28672]
    np.savez(f"{OUT}/decompose_dim.npz", n=np.array(ns), dims=np.array([get_decompose_dim(v) for v in ns]))

    # ---- a8: pack/unpack KATs from deploy/functional/quantization.py
    torch.manual_seed(1)
    q = torch.randint(-8, 8, (16, 64), dtype=torch.int8)
    q[0, :16] = torch.arange(-8, 8, dtype=torch.int8)
    p = dq.pack_i4(q)
    u = dq.unpack_i4(p)
    assert torch.equal(u, q.int())
    np.savez(f"{OUT}/pack_i4.npz", q=q.numpy(), packed=p.numpy(), unpacked=u.numpy())

    # ---- a2: kronecker_matmul on its own (incl. non power-of-two factors 14x16)
    torch.manual_seed(2)
    for (n1, n2) in [(16, 16), (14, 16), (64, 64)]:
        x = torch.randn(3, 5, n1 * n2)
        L, R = torch.randn(n1, n1), torch.randn(n2, n2)
        np.savez_compressed(f"{OUT}/kron_{n1}x{n2}.npz", x=x.numpy(), left=L.numpy(), right=R.numpy(),
                            z=kronecker_matmul(x, L, R).numpy())

    # ---- a4: ActivationQuantizer (LAC, clip_ratio, plain; zero rows; fp16 arithmetic)
    torch.manual_seed(3)
    x = torch.randn(12, 256)
    x[3] = 0
    x[4] = x[4].abs()  # all-positive row: lo clamps to 0
    cases = {}
    for name, kw, ca in [("lac", dict(lac=True), (4.0, 2.0)), ("plain", dict(), None),
                         ("ratio", dict(clip_ratio=0.9), None)]:
        aq = ActivationQuantizer(bits=4, sym=True, **kw)
        if ca:
            with torch.no_grad():
                aq.clip_factor_a_max.fill_(ca[0])
                aq.clip_factor_a_min.fill_(ca[1])
        with torch.no_grad():
            s, _ = aq.get_scale_zero(x)
            cases[f"{name}_scale"] = s[:, :1].numpy()
            cases[f"{name}_fq"] = aq(x).numpy()
            xh = x.half()
            sh, _ = aq.get_scale_zero(xh)
            cases[f"{name}_scale_h"] = sh[:, :1].numpy()
            cases[f"{name}_fq_h"] = aq(xh).numpy()
    np.savez_compressed(f"{OUT}/act_quant.npz", x=x.numpy(), **cases)

    # ---- a4 (vLLM copy): grouped quantisation and caller-supplied scale/zero (vllm_custom/.../fake_quant_utils.py:69-88)
    vq = load_by_path("ref_vllm_fake_quant",
                      f"{REF}/vllm_custom/model_executor/layers/quantization/utils/fake_quant_utils.py")
    torch.manual_seed(31)
    xv = torch.randn(5, 3, 128)
    vcases = {"x": xv.numpy()}
    with torch.no_grad():
        gq = vq.ActivationQuantizer(bits=4, sym=False, lac=True, groupsize=64)
        gq.clip_factor_a_max.fill_(3.0)
        gq.clip_factor_a_min.fill_(5.0)
        vcases["grouped_asym_lac"] = gq(xv).numpy()
        sq = vq.ActivationQuantizer(bits=4, sym=True)
        s_given, z_given = sq.get_scale_zero(xv * 1.5)
        vcases["given_scale"], vcases["given_zero"] = s_given.numpy(), z_given.numpy()
        vcases["sym_given"] = sq(xv, s_given, z_given).numpy()
    np.savez_compressed(f"{OUT}/act_quant_vllm.npz", **vcases)

    # ---- a5: WeightQuantizer RTN
    torch.manual_seed(4)
    w = torch.randn(32, 128) * torch.rand(32, 1)
    w[5] = 0
    wq = WeightQuantizer()
    wq.configure(4, perchannel=True, sym=True, mse=False)
    wq.find_params(w)
    np.savez_compressed(f"{OUT}/weight_rtn.npz", w=w.numpy(), scale=wq.scale.numpy(), w_fq=wq.quantize(w).numpy())

    # ---- a10: deploy sym_quant (torch reference in deploy/functional/quantization.py, fp16 math)
    torch.manual_seed(5)
    xh = (torch.randn(8, 128) * 3).half()
    sh = (xh.abs().amax(1, keepdim=True) / 7).half()
    qh, _ = dq.sym_quant(xh, sh, torch.tensor(7))
    np.savez_compressed(f"{OUT}/sym_quant_half.npz", x=xh.numpy(), scale=sh.numpy(), q=qh.numpy())

    # ---- a6/a7: the whole FlatQuantizedLinear path (cfg1 family)
    flat_linear_case("flatlinear_256_svd", m=8, k=256, n=64, seed=0, trans_kind="svd")
    flat_linear_case("flatlinear_224_inv_bias", m=6, k=224, n=48, seed=1, trans_kind="inv", bias=True, ca=(3.0, 5.0))
    flat_linear_case("flatlinear_4096_svd", m=16, k=4096, n=64, seed=0, trans_kind="svd")

    # ---- o_proj style: single matrix over the heads axis + list-form qa_trans (llama_utils.py:264-283)
    torch.manual_seed(6)
    np.random.seed(6)
    heads, hd, n = 8, 16, 32
    o_tr = SVDSingleTransMatrix(heads)
    v_tr = SVDSingleTransMatrix(hd)
    o_tr.to_eval_mode()
    v_tr.to_eval_mode()
    lin = torch.nn.Linear(heads * hd, n, bias=False)
    w0 = lin.weight.data.clone()
    fl = FlatQuantizedLinear(args_ns(), lin)
    attn = torch.randn(2, 5, heads * hd)
    with torch.no_grad():
        fl.reparameterize(qa_trans=[o_tr.get_matrix(inv_t=True), v_tr.get_matrix(inv_t=True)])
        w_rep = lin.weight.data.clone()
        wq_mod = rtn(lin)
        a = attn.reshape(-1, heads, hd)
        z = torch.matmul(o_tr.get_matrix().T.to(a), a).reshape(attn.shape)
        y = fl(z)
    np.savez_compressed(
        f"{OUT}/flatlinear_oproj.npz", attn=attn.numpy(), w0=w0.numpy(), o=o_tr.matrix.data.numpy(),
        o_inv_t=o_tr.matrix_inv_t.data.numpy(), v_inv_t=v_tr.matrix_inv_t.data.numpy(),
        cw_max=fl.clip_factor_w_max.data.numpy(), cw_min=fl.clip_factor_w_min.data.numpy(),
        w_rep=w_rep.numpy(), s_w=wq_mod.scale.reshape(-1, 1).numpy(), w_fq=lin.weight.data.numpy(),
        z=z.numpy(), y=y.numpy())

    # ---- a14: KV asym int4 quant/dequant (kv_cache.py functions, extracted)
    ns_kv = {"torch": torch, "get_minq_maxq": dq.get_minq_maxq}
    extract_functions(f"{REF}/deploy/transformers/kv_cache.py",
                      {"asym_quantize_and_pack_i4", "unpack_i4_and_asym_dequantize"}, ns_kv)
    torch.manual_seed(7)
    kv = torch.randn(2, 3, 4, 128).half()
    kv[0, 0, 0] = 0.25  # constant row: (max-min) clamps to 1e-5
    one = torch.tensor(1.0)
    packed, s, zero = ns_kv["asym_quantize_and_pack_i4"](kv, one, one)
    deq = ns_kv["unpack_i4_and_asym_dequantize"](packed, s, zero)
    np.savez_compressed(f"{OUT}/kv_asym.npz", x=kv.numpy(), packed=packed.numpy(), scale=s.numpy(),
                        zero=zero.numpy(), deq=deq.numpy())

    # ---- Hadamard (flatquant/hadamard_utils.py:matmul_hadU, pure-torch butterflies)
    sys.modules["fast_hadamard_transform"] = types.ModuleType("fast_hadamard_transform")
    sys.modules["flatquant.utils"] = types.ModuleType("flatquant.utils")  # needs accelerate otherwise
    sys.modules["flatquant.utils"].cleanup_memory = lambda *a, **k: None
    hu = load_by_path("ref_hadamard_utils", f"{REF}/flatquant/hadamard_utils.py")
    torch.manual_seed(8)
    had = {}
    for n in (32, 128, 1024, 4096):
        xx = torch.randn(5, n)
        had[f"x_{n}"] = xx.numpy()
        had[f"y_{n}"] = hu.matmul_hadU(xx).numpy()  # x @ H_n / sqrt(n)
    np.savez_compressed(f"{OUT}/hadamard.npz", **had)
    print("done")


if __name__ == "__main__":
    main()
