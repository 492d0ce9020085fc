"""Golden fixture for SURVEY.md section 8(f) rows 1 and 3: a whole W4A4 Llama decoder stack and its on-disk formats,
produced by running the *reference's own code* on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_llama.py

What runs unmodified from /root/reference:
  flatquant/model_tools/llama_utils.py   FlatQuantLlamaAttention / FlatQuantLlamaMLP (constructors, reparameterize,
                                         forward = the fake-quant oracle of the north star)
  flatquant/flat_utils.py                reparameterize_model, save_flat_matrices, save_quantized_weights_with_safetensors
  gptq_utils.py                          rtn_fwrd (RTN weight quantisation)
  deploy/functional/quantization.py      pack_i4 (loaded by path; deploy/__init__ needs the unbuilt _CUDA module)

Shims (this image has transformers 5.5, the reference was written against 4.3x; `accelerate` is absent):
  * `accelerate` / `accelerate.utils` are empty stub modules (flatquant/utils.py imports names from them at module
    load; none is called here)
  * `deploy` / `deploy.functional` are stub packages exposing the by-path loaded quantization module
  * transformers 5.5's LlamaAttention no longer sets `num_heads`, `num_key_value_heads`, `hidden_size`, `rotary_emb`;
    they are set on the instance after construction (`rotary_emb(x, seq_len=)` -> LlamaRotaryEmbedding(x, arange));
    llama_utils.py:218 calls `apply_rotary_pos_emb(q, k, cos, sin, position_ids)`, whose 5th positional argument is
    now `unsqueeze_dim`, so the forward is called with `position_ids=1` (= the head axis of [b, h, s, d])
  * the decoder layer / model glue (residuals, RMSNorm, embedding, final norm, lm_head) is composed by hand because
    the HF 5.5 decoder layer calls its attention with a different signature; it is the standard pre-norm Llama layer
    the reference inherits unchanged from HF.

Outputs:
  tests/golden/llama_ckpt/{model.safetensors, quantization_config.json, flat_matrices.pth, config.json}
  tests/golden/llama_model.npz   input ids, per-stage expected activations and the final logits

`--full-width` writes the same for ONE decoder layer at the Llama-3-8B widths (hidden 4096 -> (64,64), intermediate
14336 -> (112,128), 32 heads / 8 kv heads x 128; 1 x 16 tokens):
  tests/golden/llama_fullwidth_ckpt/   (110 MB of packed weights: git-ignored, travels to the GPU box with the working
                                        tree like the built libraries; __graft_entry__.build() regenerates it when absent)
  tests/golden/llama_fullwidth.npz     (committed, 3 MB: ids + per-stage expected activations in fp32 -- they are also the
                                        teacher-forced INPUTS of the later stages, where an fp16 rounding would flip codes)
"""
import importlib.util
import json
import os
import shutil
import sys
import types

import numpy as np
import torch
import transformers  # noqa: F401  (must be imported before `accelerate` is stubbed)
from transformers import LlamaConfig, LlamaForCausalLM
from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(OUT, "llama_ckpt")
sys.path.insert(0, REF)

for name, attrs in (("accelerate", ("dispatch_model", "infer_auto_device_map")),
                    ("accelerate.utils", ("get_balanced_memory",))):
    mod = types.ModuleType(name)
    for a in attrs:
        setattr(mod, a, None)
    sys.modules[name] = mod


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


dq = load_by_path("ref_deploy_quantization", f"{REF}/deploy/functional/quantization.py")
deploy_stub = types.ModuleType("deploy")
deploy_stub.__path__ = []
deploy_fn = types.ModuleType("deploy.functional")
deploy_fn.pack_i4 = dq.pack_i4
deploy_stub.functional = deploy_fn
sys.modules["deploy"], sys.modules["deploy.functional"] = deploy_stub, deploy_fn

import gptq_utils  # noqa: E402
from flatquant import flat_utils  # noqa: E402
from flatquant.model_tools import llama_utils as LU  # noqa: E402

CFG = dict(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=2,
           num_hidden_layers=2, vocab_size=64, max_position_embeddings=128, rms_norm_eps=1e-5, rope_theta=10000.0,
           attention_bias=False, mlp_bias=False, tie_word_embeddings=False)
ARGS = dict(w_bits=4, a_bits=4, q_bits=16, k_bits=4, v_bits=4, w_asym=False, a_asym=False, q_asym=False, k_asym=True,
            v_asym=True, lac=True, lwc=True, a_groupsize=-1, w_groupsize=-1, direct_inv=False, add_diag=True,
            separate_vtrans=False, diag_init=None, gptq_mse=False, model="tiny-llama-flatquant-fixture")
BSZ, SEQ = 2, 16
FULL = "--full-width" in sys.argv
if FULL:
    CFG.update(hidden_size=4096, intermediate_size=14336, num_attention_heads=32, num_key_value_heads=8, num_hidden_layers=1)
    ARGS.update(model="llama3-8b-width-one-layer-fixture")
    CKPT = os.path.join(OUT, "llama_fullwidth_ckpt")
    BSZ = 1


class Rotary:
    """`rotary_emb(x, seq_len=)` of transformers 4.3x on top of the 5.5 module."""

    def __init__(self, cfg):
        self.inner = LlamaRotaryEmbedding(cfg)

    def __call__(self, x, seq_len=None):
        return self.inner(x, torch.arange(seq_len)[None])


def jitter_(t, lo, hi):
    with torch.no_grad():
        t.copy_(lo + (hi - lo) * torch.rand_like(t))


def main():
    torch.manual_seed(1234)
    np.random.seed(1234)
    cfg = LlamaConfig(**CFG)
    cfg._attn_implementation = "eager"
    args = types.SimpleNamespace(**ARGS)
    model = LlamaForCausalLM(cfg).float()
    with torch.no_grad():
        for p in model.parameters():  # HF init is N(0, 0.02): too flat to exercise the quantisers
            if p.dim() == 2:
                p.normal_(0, 1.0 / p.shape[1] ** 0.5)
        model.model.embed_tokens.weight.normal_(0, 1.0)
    for layer in model.model.layers:
        jitter_(layer.input_layernorm.weight, 0.6, 1.4)
        jitter_(layer.post_attention_layernorm.weight, 0.6, 1.4)
        layer.self_attn = LU.FlatQuantLlamaAttention(args, layer.self_attn)
        layer.mlp = LU.FlatQuantLlamaMLP(args, layer.mlp)
        at, mlp = layer.self_attn, layer.mlp
        at.num_heads, at.num_key_value_heads, at.hidden_size = (cfg.num_attention_heads, cfg.num_key_value_heads,
                                                                 cfg.hidden_size)
        at.rotary_emb = Rotary(cfg)
        # "calibrated" parameters: non-trivial singular values, diagonal scales and clip factors
        for tr in (at.ln_trans, mlp.up_gate_trans, mlp.down_trans):
            jitter_(tr.linear_diag_left, 0.8, 1.25)
            jitter_(tr.linear_diag_right, 0.8, 1.25)
            jitter_(tr.diag_scale, 0.7, 1.3)
        for tr in (at.o_trans, at.kcache_trans, at.vcache_trans):
            jitter_(tr.linear_diag, 0.8, 1.25)
        for fl in (at.q_proj, at.k_proj, at.v_proj, at.o_proj, mlp.up_proj, mlp.gate_proj, mlp.down_proj):
            jitter_(fl.act_quantizer.clip_factor_a_max, 2.5, 5.0)
            jitter_(fl.act_quantizer.clip_factor_a_min, 2.5, 5.0)
            jitter_(fl.clip_factor_w_max, 3.0, 5.0)
            jitter_(fl.clip_factor_w_min, 3.0, 5.0)
        for q in (at.k_cache_quantizer, at.v_cache_quantizer):
            jitter_(q.clip_factor_a_max, 3.0, 5.0)
            jitter_(q.clip_factor_a_min, 3.0, 5.0)
        # gate and up share the quantised activation in the deployed model only if their clip factors agree; keep one
        # layer with equal factors (fused gate/up GEMM) and one with distinct factors (separate quantisations)
    l0 = model.model.layers[0].mlp
    with torch.no_grad():
        l0.gate_proj.act_quantizer.clip_factor_a_max.copy_(l0.up_proj.act_quantizer.clip_factor_a_max)
        l0.gate_proj.act_quantizer.clip_factor_a_min.copy_(l0.up_proj.act_quantizer.clip_factor_a_min)

    if os.path.isdir(CKPT):
        shutil.rmtree(CKPT)
    os.makedirs(CKPT)
    args.exp_dir = CKPT
    with torch.no_grad():
        flat_utils.save_flat_matrices(args, model)          # flat_matrices.pth (eval-mode matrices + clip factors)
        flat_utils.reparameterize_model(model)              # bakes transforms into the weights / RMSNorm scales
        quantizers = gptq_utils.rtn_fwrd(model, "cpu", args)
        flat_utils.save_quantized_weights_with_safetensors(args, model, quantizers)
    with open(os.path.join(CKPT, "config.json"), "w") as f:
        json.dump(dict(CFG, model_type="llama_FlatQuant", architectures=["FlatQuantLlamaForCausalLM"],
                       quantization_config=dict(fuseLN=False, trans="matmul",
                                                online_trans=["qk", "o_proj", "down_proj", "qkv_proj", "up_gate_proj"])),
                  f, indent=1)

    # The export stores everything but the packed weights and their scales in fp16.  Round the live model's parameters
    # the same way so that the forward below runs on exactly the values a loader of the checkpoint sees (W4A4 code
    # flips make a deep stack sensitive to 1e-3 parameter perturbations; this keeps the fixture about arithmetic).
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if not name.endswith("linear.weight"):
                prm.copy_(prm.half().float())

    # ---- the fake-quant forward of the reference, stage by stage
    torch.manual_seed(99)
    ids = torch.randint(0, cfg.vocab_size, (BSZ, SEQ))
    mask = torch.full((SEQ, SEQ), float("-inf")).triu(1)[None, None].expand(BSZ, 1, SEQ, SEQ)
    rec = {}

    def run(tag, kv_quant):
        caps = {}
        hooks = []
        for li, layer in enumerate(model.model.layers):
            at = layer.self_attn
            at.k_cache_quantizer.enable = kv_quant
            at.v_cache_quantizer.enable = kv_quant
            for mod, names in ((at, ("q_proj", "k_proj", "v_proj", "o_proj")), (layer.mlp, ("gate_proj", "up_proj"))):
                for nm in names:
                    hooks.append(getattr(mod, nm).register_forward_hook(
                        lambda m, i, o, key=f"l{li}_{nm}": caps.__setitem__(key, o.detach().clone())))
        with torch.no_grad():
            h = model.model.embed_tokens(ids)
            for li, layer in enumerate(model.model.layers):
                a = layer.self_attn(layer.input_layernorm(h), mask, 1, None, False, False)[0]
                h = h + a
                caps[f"l{li}_attn"] = a.clone()
                m = layer.mlp(layer.post_attention_layernorm(h))
                h = h + m
                caps[f"l{li}_mlp"] = m.clone()
                caps[f"l{li}_out"] = h.clone()
            logits = model.lm_head(model.model.norm(h))
        for hk in hooks:
            hk.remove()
        if tag == "prefill":  # stage-by-stage captures for teacher-forced checks of every quantised linear
            for k, v in caps.items():
                rec[f"{tag}_{k}"] = v.numpy()
        rec[f"{tag}_logits"] = logits.numpy()
        print(tag, "logits absmax", float(logits.abs().max()), "layer-1 out absmax", float(h.abs().max()))

    run("prefill", kv_quant=False)   # = the deployed prefill: attention on the transformed, unquantised K/V
    run("kvq", kv_quant=True)        # fake-quantised (LAC, asymmetric INT4) K/V in the attention
    np.savez_compressed(os.path.join(OUT, "llama_fullwidth.npz" if FULL else "llama_model.npz"), ids=ids.numpy(), **rec)
    print("files:", {f: os.path.getsize(os.path.join(CKPT, f)) for f in sorted(os.listdir(CKPT))})


if __name__ == "__main__":
    main()
