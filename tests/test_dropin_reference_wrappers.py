"""SURVEY.md section 8(b), row "Python operator API": the reference's OWN model wrappers, unmodified, running on this
package's operator modules.

`flatquant/model_tools/llama_utils.py` (FlatQuantLlamaAttention / FlatQuantLlamaMLP) is loaded twice from
/root/reference: once as it is (importing the reference's flatquant.{flat_linear, quant_utils, trans_utils,
function_utils}), once with those four module names resolving to `flatquant_b200.*`.  Same seeds, same tiny Llama
layer, then: identical state-dict keys and initial parameters, identical calibration-mode forward (the fake-quant
path), identical weights after `reparameterize()` and RTN, a state dict saved by one loads into the other, and the
reference's `flat_utils.reparameterize_model` / `save_flat_matrices` / `gptq_utils.rtn_fwrd` accept the model built
on our modules.  The eval-mode forward of our FlatQuantizedLinear is CUDA-only; it is compared with the reference's
outputs on the GPU in tests/test_gpu_parity.py::test_flat_quantized_linear_dropin and tests/test_zz_llama_model_gpu.py.

CPU-only and only in the build container: the reference cannot travel to the GPU box, so the test skips itself when
/root/reference is absent.  Shims for transformers 5.5 are the ones listed in tests/golden/make_golden_llama.py."""
import copy
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "flatquant")), reason="needs /root/reference")

SWAPPED = ("flat_linear", "quant_utils", "trans_utils", "function_utils")
CFG = dict(hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=1, num_hidden_layers=1,
           vocab_size=32, max_position_embeddings=64, rms_norm_eps=1e-5, rope_theta=10000.0, attention_bias=False,
           mlp_bias=False, tie_word_embeddings=False)
ARGS = dict(w_bits=4, a_bits=4, q_bits=16, k_bits=4, v_bits=4, w_asym=False, a_asym=False, q_asym=False, k_asym=True,
            v_asym=True, lac=True, lwc=True, a_groupsize=-1, w_groupsize=-1, direct_inv=False, add_diag=True,
            separate_vtrans=False, diag_init=None, gptq_mse=False, model="tiny-llama-dropin")
BSZ, SEQ = 2, 10


def _stub_missing_modules():
    import transformers  # noqa: F401  (before `accelerate` is stubbed: transformers probes for it at import)
    for name, attrs in (("accelerate", ("dispatch_model", "infer_auto_device_map")),
                        ("accelerate.utils", ("get_balanced_memory",))):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for a in attrs:
                setattr(mod, a, None)
            sys.modules[name] = mod


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def wrappers():
    """(reference llama_utils on the reference's modules, the same file on flatquant_b200's modules)"""
    _stub_missing_modules()
    added = REF not in sys.path
    if added:
        sys.path.insert(0, REF)
    try:
        import flatquant  # the reference package (this repo's package is `flatquant_b200`: no name clash)
        assert os.path.realpath(os.path.dirname(flatquant.__file__)).startswith(REF)
        path = os.path.join(REF, "flatquant", "model_tools", "llama_utils.py")
        ref = _load_by_path("ref_llama_utils_on_reference_ops", path)
        import flatquant_b200
        saved = {n: sys.modules.get(f"flatquant.{n}") for n in SWAPPED}
        try:
            for n in SWAPPED:
                sys.modules[f"flatquant.{n}"] = importlib.import_module(f"flatquant_b200.{n}")
            ours = _load_by_path("ref_llama_utils_on_b200_ops", path)
        finally:
            for n, m in saved.items():
                if m is None:
                    sys.modules.pop(f"flatquant.{n}", None)
                else:
                    sys.modules[f"flatquant.{n}"] = m
        assert ours.FlatQuantizedLinear is flatquant_b200.flat_linear.FlatQuantizedLinear
        assert ref.FlatQuantizedLinear is not ours.FlatQuantizedLinear
        yield ref, ours
    finally:
        if added:
            sys.path.remove(REF)


def _tiny_layer():
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaAttention, LlamaMLP
    cfg = LlamaConfig(**CFG)
    cfg._attn_implementation = "eager"
    torch.manual_seed(7)
    attn, mlp = LlamaAttention(cfg, 0).float(), LlamaMLP(cfg).float()
    with torch.no_grad():
        for p in list(attn.parameters()) + list(mlp.parameters()):
            p.normal_(0, 1.0 / p.shape[1] ** 0.5)
    return cfg, attn, mlp


class _Rotary:
    """`rotary_emb(x, seq_len=)` of transformers 4.3x on top of the 5.5 module (see make_golden_llama.py)."""

    def __init__(self, cfg):
        from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding
        self.inner = LlamaRotaryEmbedding(cfg)

    def __call__(self, x, seq_len=None):
        return self.inner(x, torch.arange(seq_len)[None])


def _wrap(lu, cfg, attn, mlp):
    """Wrap copies of (attn, mlp) with the wrappers of `lu`, from a fixed RNG state (the transforms draw their random
    orthogonal factors from numpy's global RNG, function_utils.py:24-28), then jitter every learnable factor."""
    args = types.SimpleNamespace(**ARGS)
    torch.manual_seed(11)
    np.random.seed(11)
    at = lu.FlatQuantLlamaAttention(args, copy.deepcopy(attn))
    ml = lu.FlatQuantLlamaMLP(args, copy.deepcopy(mlp))
    at.num_heads, at.num_key_value_heads, at.hidden_size = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.hidden_size
    at.rotary_emb = _Rotary(cfg)
    gen = torch.Generator().manual_seed(5)

    def jitter_(t, lo, hi):
        with torch.no_grad():
            t.copy_(lo + (hi - lo) * torch.rand(t.shape, generator=gen))
    for tr in (at.ln_trans, ml.up_gate_trans, ml.down_trans):
        jitter_(tr.linear_diag_left, 0.8, 1.25), jitter_(tr.linear_diag_right, 0.8, 1.25), jitter_(tr.diag_scale, 0.7, 1.3)
    for tr in (at.o_trans, at.kcache_trans, at.vcache_trans):
        jitter_(tr.linear_diag, 0.8, 1.25)
    for fl in (at.q_proj, at.k_proj, at.v_proj, at.o_proj, ml.up_proj, ml.gate_proj, ml.down_proj):
        jitter_(fl.act_quantizer.clip_factor_a_max, 2.5, 5.0), jitter_(fl.act_quantizer.clip_factor_a_min, 2.5, 5.0)
        jitter_(fl.clip_factor_w_max, 3.0, 5.0), jitter_(fl.clip_factor_w_min, 3.0, 5.0)
    for q in (at.k_cache_quantizer, at.v_cache_quantizer):
        jitter_(q.clip_factor_a_max, 3.0, 5.0), jitter_(q.clip_factor_a_min, 3.0, 5.0)
    return args, at, ml


def _forward(at, ml, x):
    mask = torch.full((SEQ, SEQ), float("-inf")).triu(1)[None, None].expand(BSZ, 1, SEQ, SEQ)
    with torch.no_grad():
        a = at(x, mask, 1, None, False, False)[0]  # position_ids=1: see make_golden_llama.py (unsqueeze_dim)
        return a, ml(x + a)


def _same_state(a, b, exact=True):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype, k
        if exact:
            assert torch.equal(sa[k], sb[k]), k
        else:
            assert torch.allclose(sa[k], sb[k], rtol=1e-6, atol=1e-7), k


def test_reference_wrappers_run_unchanged_on_our_operator_modules(wrappers):
    ref_lu, our_lu = wrappers
    cfg, attn, mlp = _tiny_layer()
    _, r_at, r_ml = _wrap(ref_lu, cfg, attn, mlp)
    _, o_at, o_ml = _wrap(our_lu, cfg, attn, mlp)
    from flatquant_b200.flat_linear import FlatQuantizedLinear
    from flatquant_b200.trans_utils import SVDDecomposeTransMatrix, SVDSingleTransMatrix
    assert isinstance(o_at.q_proj, FlatQuantizedLinear) and isinstance(o_ml.down_trans, SVDDecomposeTransMatrix)
    assert isinstance(o_at.o_trans, SVDSingleTransMatrix) and not isinstance(r_at.q_proj, FlatQuantizedLinear)
    # (1) constructors: same parameter / buffer names, shapes, dtypes and initial values (state-dict contract)
    _same_state(r_at, o_at)
    _same_state(r_ml, o_ml)
    # (2) calibration-mode forward = the fake-quant path (transform, LAC/LWC clipping, RTN, F.linear)
    torch.manual_seed(3)
    x = torch.randn(BSZ, SEQ, cfg.hidden_size)
    (ra, rm), (oa, om) = _forward(r_at, r_ml, x), _forward(o_at, o_ml, x)
    assert torch.allclose(oa, ra, rtol=1e-5, atol=1e-6 * float(ra.abs().max()))
    assert torch.allclose(om, rm, rtol=1e-5, atol=1e-6 * float(rm.abs().max()))
    # ... with gradients flowing to the same learnable parameters (calibration trains them)
    for at_, ml_ in ((r_at, r_ml), (o_at, o_ml)):
        mask = torch.full((SEQ, SEQ), float("-inf")).triu(1)[None, None].expand(BSZ, 1, SEQ, SEQ)
        a = at_(x, mask, 1, None, False, False)[0]
        (ml_(x + a).pow(2).mean() + a.pow(2).mean()).backward()
    rg = {k: p.grad for k, p in list(r_at.named_parameters()) + list(r_ml.named_parameters())}
    og = {k: p.grad for k, p in list(o_at.named_parameters()) + list(o_ml.named_parameters())}
    assert {k for k, g in rg.items() if g is not None} == {k for k, g in og.items() if g is not None}
    for k, g in rg.items():
        if g is not None:
            assert torch.allclose(og[k], g, rtol=1e-3, atol=1e-5 * float(g.abs().max()) + 1e-12), k
    # (3) a state dict written by the reference modules loads into ours and vice versa
    o_at.load_state_dict(r_at.state_dict())
    r_ml.load_state_dict(o_ml.state_dict())
    # (4) reparameterize(): transforms to eval mode, weights baked in fp64, diag scale folded into up_proj
    with torch.no_grad():
        for m in (r_at, r_ml, o_at, o_ml):
            m.reparameterize()
    _same_state(r_at, o_at, exact=False)
    _same_state(r_ml, o_ml, exact=False)
    assert o_at.q_proj._eval_mode and o_at._eval_mode and o_ml.up_gate_trans.use_diag is False
    # eval-mode transforms (the online part of the hot path) agree on CPU
    with torch.no_grad():
        assert torch.allclose(o_at.ln_trans(x), r_at.ln_trans(x), rtol=1e-5, atol=1e-6)
        assert torch.allclose(o_at.kcache_trans(x.view(BSZ, SEQ, 2, 128), inv_t=True),
                              r_at.kcache_trans(x.view(BSZ, SEQ, 2, 128), inv_t=True), rtol=1e-5, atol=1e-6)
    # (5) our eval forward is CUDA-only: it must refuse CPU activations loudly, not fall back
    with pytest.raises(RuntimeError, match="CUDA-only"):
        o_at.q_proj(x)


def test_reference_model_level_tools_accept_our_modules(wrappers, tmp_path):
    """flat_utils.reparameterize_model, flat_utils.save_flat_matrices and gptq_utils.rtn_fwrd of the reference, called
    on a model whose quantised layers are built from our modules, produce what they produce on the reference's."""
    ref_lu, our_lu = wrappers
    sys.path.insert(0, REF)
    try:
        dq = _load_by_path("ref_deploy_quantization", f"{REF}/deploy/functional/quantization.py")
        if "deploy" not in sys.modules:  # gptq_utils / flat_utils import deploy lazily for the safetensors export only
            stub, fn = types.ModuleType("deploy"), types.ModuleType("deploy.functional")
            stub.__path__, fn.pack_i4, stub.functional = [], dq.pack_i4, fn
            sys.modules["deploy"], sys.modules["deploy.functional"] = stub, fn
        import gptq_utils
        from flatquant import flat_utils as ref_flat_utils
    finally:
        sys.path.remove(REF)
    cfg, attn, mlp = _tiny_layer()
    results = []
    for lu in (ref_lu, our_lu):
        args, at, ml = _wrap(lu, cfg, attn, mlp)
        layer = torch.nn.Module()
        layer.self_attn, layer.mlp = at, ml
        layer.input_layernorm = torch.nn.LayerNorm(cfg.hidden_size, elementwise_affine=True, bias=False)
        layer.post_attention_layernorm = torch.nn.LayerNorm(cfg.hidden_size, elementwise_affine=True, bias=False)
        model = torch.nn.Module()
        model.model = torch.nn.Module()
        model.model.layers = torch.nn.ModuleList([layer])
        model.config = cfg
        args.exp_dir = str(tmp_path / lu.__name__)
        os.makedirs(args.exp_dir)
        with torch.no_grad():
            ref_flat_utils.save_flat_matrices(args, model)
            ref_flat_utils.reparameterize_model(model)
            quantizers = gptq_utils.rtn_fwrd(model, "cpu", args)
        results.append((model, quantizers, torch.load(os.path.join(args.exp_dir, "flat_matrices.pth"))))
    (rm, rq, rf), (om, oq, of) = results
    assert list(rf.keys()) == list(of.keys()) and list(rf[0].keys()) == list(of[0].keys())
    for k in rf[0]:
        assert torch.allclose(rf[0][k], of[0][k], rtol=1e-6, atol=1e-7), k
    assert list(rq.keys()) == list(oq.keys())
    for k in rq:  # per-channel weight scales found by RTN on the reparameterised weights
        assert torch.allclose(rq[k].scale, oq[k].scale, rtol=1e-5), k
    rs, os_ = rm.state_dict(), om.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    for k in rs:
        if k.endswith("linear.weight"):  # on the INT4 grid: equal codes except where fp32 rounding straddles a boundary
            d = (rs[k] - os_[k]).abs()
            assert float((d > 1e-6 * float(rs[k].abs().max())).float().mean()) < 1e-3, k
        else:
            assert torch.allclose(rs[k], os_[k], rtol=1e-5, atol=1e-6), k
    # the RTN-rounded weights of the model built on our modules pack through our own pack_weights (INT4 grid check)
    w_packed, s_w = om.model.layers[0].mlp.down_proj.pack_weights()
    assert w_packed.dtype == torch.uint8 and w_packed.shape == (cfg.hidden_size, cfg.intermediate_size // 2)
    assert torch.allclose(s_w, oq["model.layers.0.mlp.down_proj.linear"].scale.reshape(-1).float(), rtol=1e-4)


def test_reference_deepseek_wrappers_on_our_modules_reproduce_the_fixture():
    """flatquant/model_tools/deepseekv3_utils.py (FlatQuantMLA, FlatQuantMoE and its own FlatQuantizedLinear subclass
    for the DeepSeek linears) imports WeightQuantizer / ActivationQuantizer, kronecker_matmul, the SVD transform
    classes and get_decompose_dim from the flatquant package.  Loaded with those names resolving to flatquant_b200.*,
    the block built by the fixture generator (same seeds) reproduces the committed reference outputs."""
    from conftest import GOLDEN
    _stub_missing_modules()
    sys.path.insert(0, REF)
    try:
        gen = _load_by_path("make_golden_deepseek_for_dropin", os.path.join(GOLDEN, "make_golden_deepseek.py"))
        names = ("quant_utils", "flat_utils", "trans_utils", "function_utils")
        saved = {n: sys.modules.get(f"flatquant.{n}") for n in names}
        try:
            for n in names:
                sys.modules[f"flatquant.{n}"] = importlib.import_module(f"flatquant_b200.{n}")
            du = _load_by_path("ref_deepseekv3_utils_on_b200_ops",
                               os.path.join(REF, "flatquant", "model_tools", "deepseekv3_utils.py"))
        finally:
            for n, m in saved.items():
                if m is None:
                    sys.modules.pop(f"flatquant.{n}", None)
                else:
                    sys.modules[f"flatquant.{n}"] = m
    finally:
        sys.path.remove(REF)
    import flatquant_b200.quant_utils as our_q
    import flatquant_b200.trans_utils as our_t
    assert du.ActivationQuantizer is our_q.ActivationQuantizer and du.SVDDecomposeTransMatrix is our_t.SVDDecomposeTransMatrix
    block, args, transforms, linears = gen.build_block(du)
    assert isinstance(block.attn.qkv_trans, our_t.SVDDecomposeTransMatrix)
    assert isinstance(block.attn.wq_a.weight_quantizer, our_q.WeightQuantizer)
    x, freqs, mask = gen.fixture_input(args)
    g = np.load(os.path.join(GOLDEN, "deepseek_block.npz"))
    assert np.array_equal(x.numpy(), g["x"])
    with torch.no_grad():
        y = block(x, 0, freqs, mask)
    err = np.abs(y.numpy() - g["y"]).max() / np.abs(g["y"]).max()
    assert err < 1e-5, err
    for name, tr in transforms.items():  # eval-mode matrices equal the ones the reference modules produced
        assert np.allclose(tr.matrix_left.data.numpy(), g[f"{name}.matrix_left"], rtol=1e-5, atol=1e-6), name
        assert np.allclose(tr.matrix_right.data.numpy(), g[f"{name}.matrix_right"], rtol=1e-5, atol=1e-6), name
    with torch.no_grad():  # the offline weight preparation of every DeepSeek linear gives the fixture's scales
        for name in ("wq_a", "wo", "shared.w2", "experts.3.w1", "experts.5.w2"):
            fl, tr = linears[name]
            w = fl.apply_wclip(fl.apply_trans(fl.get_weight(), tr))
            fl.weight_quantizer.find_params(w)
            assert np.allclose(fl.weight_quantizer.scale.reshape(-1, 1).numpy(), g[f"{name}.s_w"], rtol=1e-5), name


def test_vllm_custom_utility_copies_are_interchangeable_with_ours():
    """vllm_custom/model_executor/layers/quantization/utils/{flatquant_utils, fake_quant_utils}.py are the copies the
    vLLM model files import (llama_flatquant.py:57-59).  Same constructors, state-dict keys and arithmetic as ours."""
    base = os.path.join(REF, "vllm_custom", "model_executor", "layers", "quantization", "utils")
    vt = _load_by_path("ref_vllm_flatquant_utils", os.path.join(base, "flatquant_utils.py"))
    vq = _load_by_path("ref_vllm_fake_quant_utils", os.path.join(base, "fake_quant_utils.py"))
    import flatquant_b200.flat_utils as our_f
    import flatquant_b200.function_utils as our_fn
    import flatquant_b200.quant_utils as our_q
    import flatquant_b200.trans_utils as our_t
    for n in (64, 128, 512, 1024, 3584, 4096, 8192, 14336, 28672):
        assert vt.get_decompose_dim(n) == our_fn.get_decompose_dim(n)
    torch.manual_seed(0)
    x = torch.randn(3, 5, 8 * 16)
    for cls, ctor in (("SVDDecomposeTransMatrix", (8, 16)), ("InvDecomposeTransMatrix", (8, 16)),
                      ("SVDSingleTransMatrix", (16,)), ("InvSingleTransMatrix", (16,))):
        np.random.seed(1), torch.manual_seed(1)
        theirs = getattr(vt, cls)(*ctor)
        np.random.seed(1), torch.manual_seed(1)
        mine = getattr(our_t, cls)(*ctor)
        assert list(theirs.state_dict().keys()) == list(mine.state_dict().keys()), cls
        mine.load_state_dict(theirs.state_dict())
        xin = x if len(ctor) == 2 else x.view(3, 5, 8, 16)
        with torch.no_grad():
            for inv_t in (False, True):
                assert torch.allclose(mine(xin, inv_t=inv_t), theirs(xin, inv_t=inv_t), rtol=1e-5, atol=1e-6), cls
            theirs.to_eval_mode(), mine.to_eval_mode()
            mine_keys = list(mine.state_dict().keys())
            if cls == "InvSingleTransMatrix":  # the vLLM copy deletes `linear` here; the main repo's class
                mine_keys.remove("linear.weight")  # (trans_utils.py:153-159), which ours mirrors, keeps it
            assert list(theirs.state_dict().keys()) == mine_keys, cls
            for inv_t in (False, True):
                assert torch.allclose(mine(xin, inv_t=inv_t), theirs(xin, inv_t=inv_t), rtol=1e-5, atol=1e-6), cls
    L, R = torch.randn(8, 8), torch.randn(16, 16)
    assert torch.allclose(our_f.kronecker_matmul(x, L, R), vt.kronecker_matmul(x, L, R), rtol=1e-5, atol=1e-5)
    for kw in (dict(bits=4, sym=True, lac=True), dict(bits=4, sym=False, lac=True), dict(bits=4, sym=False, groupsize=32),
               dict(bits=8, sym=True, clip_ratio=0.9)):
        theirs, mine = vq.ActivationQuantizer(**kw), our_q.ActivationQuantizer(**kw)
        assert list(theirs.state_dict().keys()) == list(mine.state_dict().keys()), kw
        assert torch.equal(mine(x), theirs(x)), kw
    w = torch.randn(24, 64)
    for kw in (dict(perchannel=True, sym=True, mse=False), dict(perchannel=True, sym=False, mse=False),
               dict(perchannel=True, sym=True, mse=True)):
        theirs, mine = vq.WeightQuantizer(), our_q.WeightQuantizer()
        theirs.configure(4, **kw), mine.configure(4, **kw)
        theirs.find_params(w), mine.find_params(w)
        assert torch.equal(mine.scale, theirs.scale) and torch.equal(mine.quantize(w), theirs.quantize(w)), kw


@pytest.mark.parametrize("family", ["llama31_utils", "qwen_utils"])
def test_other_model_families_wrappers_on_our_modules(family):
    """The Llama-3.1 and Qwen2 wrappers of flatquant/model_tools (position_embeddings calling convention; Qwen2 with
    q/k/v biases, whose v bias goes through the value-cache transform in reparameterize, flat_linear.py:95-96): same
    check as for llama_utils -- reference modules vs flatquant_b200 modules under the unmodified wrapper file."""
    from transformers import LlamaConfig, Qwen2Config
    from transformers.models.llama.modeling_llama import LlamaAttention, LlamaMLP, LlamaRotaryEmbedding
    from transformers.models.qwen2.modeling_qwen2 import Qwen2Attention, Qwen2MLP
    _stub_missing_modules()
    sys.path.insert(0, REF)
    try:
        path = os.path.join(REF, "flatquant", "model_tools", family + ".py")
        ref_lu = _load_by_path(f"ref_{family}_on_reference_ops", path)
        saved = {n: sys.modules.get(f"flatquant.{n}") for n in SWAPPED}
        try:
            for n in SWAPPED:
                sys.modules[f"flatquant.{n}"] = importlib.import_module(f"flatquant_b200.{n}")
            our_lu = _load_by_path(f"ref_{family}_on_b200_ops", path)
        finally:
            for n, m in saved.items():
                if m is None:
                    sys.modules.pop(f"flatquant.{n}", None)
                else:
                    sys.modules[f"flatquant.{n}"] = m
    finally:
        sys.path.remove(REF)
    qwen = family == "qwen_utils"
    kw = dict(CFG)
    if qwen:
        kw.pop("attention_bias"), kw.pop("mlp_bias")
        cfg = Qwen2Config(**kw)
    else:
        cfg = LlamaConfig(**kw)
    cfg._attn_implementation = "eager"
    torch.manual_seed(7)
    attn = (Qwen2Attention if qwen else LlamaAttention)(cfg, 0).float()
    mlp = (Qwen2MLP if qwen else LlamaMLP)(cfg).float()
    with torch.no_grad():
        for p in list(attn.parameters()) + list(mlp.parameters()):
            if p.dim() == 2:
                p.normal_(0, 1.0 / p.shape[1] ** 0.5)
            else:
                p.normal_(0, 0.3)  # Qwen2's q/k/v biases
    names = ("FlatQuantQwen2Attention", "FlatQuantQwen2MLP") if qwen else ("FlatQuantLlamaAttention", "FlatQuantLlamaMLP")
    torch.manual_seed(3)
    x = torch.randn(BSZ, SEQ, cfg.hidden_size)
    pos = torch.arange(SEQ)[None].expand(BSZ, SEQ)
    cos_sin = LlamaRotaryEmbedding(LlamaConfig(**CFG))(x, pos)
    mask = torch.full((SEQ, SEQ), float("-inf")).triu(1)[None, None].expand(BSZ, 1, SEQ, SEQ)
    outs = []
    for lu in (ref_lu, our_lu):
        args = types.SimpleNamespace(**ARGS)
        torch.manual_seed(11)
        np.random.seed(11)
        at = getattr(lu, names[0])(args, copy.deepcopy(attn))
        ml = getattr(lu, names[1])(args, copy.deepcopy(mlp))
        at.num_heads, at.num_key_value_heads, at.hidden_size = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.hidden_size
        gen = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for name, p in list(at.named_parameters()) + list(ml.named_parameters()):
                if "clip_factor" in name:
                    p.copy_(2.5 + 2.5 * torch.rand(p.shape, generator=gen))
                elif "linear_diag" in name or "diag_scale" in name:
                    p.copy_(0.8 + 0.45 * torch.rand(p.shape, generator=gen))
            a = at(x, mask, pos, None, False, False, position_embeddings=cos_sin)[0]
            m = ml(x + a)
            at.reparameterize(), ml.reparameterize()
        outs.append((at, ml, a, m))
    (r_at, r_ml, ra, rm), (o_at, o_ml, oa, om) = outs
    from flatquant_b200.flat_linear import FlatQuantizedLinear
    assert isinstance(o_at.v_proj, FlatQuantizedLinear) and not isinstance(r_at.v_proj, FlatQuantizedLinear)
    assert torch.allclose(oa, ra, rtol=1e-5, atol=1e-6 * float(ra.abs().max()))
    assert torch.allclose(om, rm, rtol=1e-5, atol=1e-6 * float(rm.abs().max()))
    _same_state(r_at, o_at, exact=False)   # after reparameterize: baked weights, transformed v bias (Qwen2), eval matrices
    _same_state(r_ml, o_ml, exact=False)
    if qwen:
        assert o_at.v_proj.linear.bias is not None and not torch.equal(o_at.v_proj.linear.bias, attn.v_proj.bias)


def test_get_hadK_chooses_the_reference_factorisation(monkeypatch):
    """deploy/functional/online_trans.py:21-71 (loaded with its CUDA-only imports stubbed): for every hidden / intermediate
    size of the model families the reference lists, our get_hadK picks the same K x 2^m factorisation and returns the same
    K x K matrix, entry for entry (data/hadamard_tables.npz, tools/make_hadamard_tables.py)."""
    from flatquant_b200.deploy.functional import hadamard as H

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__path__ = []
        for k, v in attrs.items():
            setattr(mod, k, v)
        monkeypatch.setitem(sys.modules, name, mod)
    stub("fast_hadamard_transform"), stub("deploy"), stub("deploy.kernels")
    stub("deploy.kernels.kron_matmul", kron_matmul=None), stub("deploy.kernels.block_matmul", block_matmul=None)
    ref = _load_by_path("ref_deploy_online_trans", os.path.join(REF, "deploy", "functional", "online_trans.py"))
    for n in (4096, 14336, 11008, 13824, 5120, 27648, 8192, 6144, 1792, 20 * 64, 36 * 8, 52 * 4, 60 * 2, 140 * 16, 156 * 2,
              12 * 4, 28 * 2, 40 * 8, 108 * 32):
        hk, k = ref.get_hadK(n)
        assert k == H.hadK_order(n), n
        ours, _ = H.get_hadK(n)
        if hk is not None:
            assert torch.equal(hk.float() @ hk.float().t(), k * torch.eye(k))
            assert torch.equal(ours, hk.float()), n   # the shipped table IS the reference's, entry for entry
            assert torch.equal(H.get_hadK(n, transpose=True)[0], ref.get_hadK(n, transpose=True)[0].float())
    for n in (64, 448, 8192):
        assert ref.get_decompose_dim(n) == H.__dict__.get("get_decompose_dim", __import__(
            "flatquant_b200.function_utils", fromlist=["get_decompose_dim"]).get_decompose_dim)(n)
