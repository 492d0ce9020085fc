"""Golden fixture for BASELINE.json configs[4] (DeepSeek-V3 single block incl. MoE expert FlatLinears), produced by
running the *reference's own code* on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_deepseek.py

Unmodified from /root/reference: deepseek_v3/model.py (ModelArgs, Block, MLA, MoE, Gate, RMSNorm, precompute_freqs_cis;
its Triton kernels are imported but never called: weights are fp32, `gemm_impl == "bf16"`), and
flatquant/model_tools/deepseekv3_utils.py (FlatQuantMLA, FlatQuantMoE, FlatQuantMoEExpert, FlatQuantizedLinear --
the fake-quant forward).  No shims.

The DeepSeek wrappers never bake their transforms into the weights (`reparameterize` only switches the matrices to
eval mode, deepseekv3_utils.py:282-290,468-469): every forward transforms, clips and RTN-quantises each weight
(`_train_forward`, :75-106).  The per-linear INT4 codes / scales stored here are produced by calling exactly those
methods of the reference modules once (`apply_trans`, `apply_wclip`, `weight_quantizer.find_params/quantize`).

Output: tests/golden/deepseek_block.npz (parameters, packed weights, stage captures, block output).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import deepseek_v3.model as M  # noqa: E402

M.Linear.dtype = torch.float32
from flatquant.model_tools import deepseekv3_utils as DU  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


dq = load_by_path("ref_deploy_quantization", f"{REF}/deploy/functional/quantization.py")

ARGS = dict(max_batch_size=2, max_seq_len=32, dim=256, inter_dim=512, moe_inter_dim=128, n_layers=2, n_dense_layers=1,
            n_heads=4, n_routed_experts=8, n_shared_experts=1, n_activated_experts=2, n_expert_groups=1,
            n_limited_groups=1, score_func="softmax", route_scale=1.0, q_lora_rank=128, kv_lora_rank=64,
            qk_nope_head_dim=32, qk_rope_head_dim=32, v_head_dim=32, original_seq_len=4096, rope_theta=10000.0)
FLAT = dict(w_bits=4, a_bits=4, w_asym=False, a_asym=False, lac=True, lwc=True, a_groupsize=-1, add_diag=True)
BSZ, SEQ = 2, 12


def jitter_(t, lo, hi):
    with torch.no_grad():
        t.copy_(lo + (hi - lo) * torch.rand_like(t))


def build_block(DU=DU):
    """The fixture's block, built from a fixed RNG state on the wrapper classes of `DU` (the reference's
    deepseekv3_utils; tests/test_dropin_reference_wrappers.py passes a copy of that file whose flatquant.* imports
    resolve to flatquant_b200.*).  Returns (block, args, transforms, linears)."""
    torch.manual_seed(2024)
    np.random.seed(2024)
    args = M.ModelArgs(**ARGS)
    flat = types.SimpleNamespace(**FLAT)
    block = M.Block(1, args)  # layer 1 >= n_dense_layers: the MoE block
    with torch.no_grad():
        for name, p in block.named_parameters():
            if p.dim() == 2:
                p.normal_(0, 1.0 / p.shape[1] ** 0.5)
            else:
                jitter_(p, 0.6, 1.4)  # RMSNorm weights
        block.ffn.gate.weight.normal_(0, 0.5)  # spread the routing scores
    block.attn = DU.FlatQuantMLA(flat, block.attn)
    block.ffn = DU.FlatQuantMoE(flat, block.ffn)
    attn, ffn = block.attn, block.ffn
    transforms = {"qkv_trans": attn.qkv_trans, "wqb_trans": attn.wqb_trans, "wo_trans": attn.wo_trans,
                  "w1_trans": ffn.w1_trans, "w2_trans": ffn.w2_trans, "routed_w2_trans": ffn.routed_w2_trans}
    for tr in transforms.values():
        jitter_(tr.linear_diag_left, 0.8, 1.25)
        jitter_(tr.linear_diag_right, 0.8, 1.25)
        jitter_(tr.diag_scale, 0.7, 1.3)
    linears = {"wq_a": (attn.wq_a, attn.qkv_trans), "wkv_a": (attn.wkv_a, attn.qkv_trans),
               "wq_b": (attn.wq_b, attn.wqb_trans), "wo": (attn.wo, attn.wo_trans),
               "shared.w1": (ffn.shared_experts.w1, ffn.w1_trans), "shared.w3": (ffn.shared_experts.w3, ffn.w1_trans),
               "shared.w2": (ffn.shared_experts.w2, ffn.w2_trans)}
    for e, ex in enumerate(ffn.experts):
        linears[f"experts.{e}.w1"] = (ex.w1, ffn.w1_trans)
        linears[f"experts.{e}.w3"] = (ex.w3, ffn.w1_trans)
        linears[f"experts.{e}.w2"] = (ex.w2, ffn.routed_w2_trans)
    quantizers = {id(fl.act_quantizer): fl.act_quantizer for fl, _ in linears.values()}
    for q in quantizers.values():
        jitter_(q.clip_factor_a_max, 2.5, 5.0)
        jitter_(q.clip_factor_a_min, 2.5, 5.0)
    for fl, _ in linears.values():
        jitter_(fl.clip_factor_w_max, 3.0, 5.0)
        jitter_(fl.clip_factor_w_min, 3.0, 5.0)
    attn.reparameterize()
    ffn.reparameterize()
    return block, args, transforms, linears


def fixture_input(args):
    torch.manual_seed(7)
    x = torch.randn(BSZ, SEQ, args.dim)
    freqs = M.precompute_freqs_cis(args)[:SEQ]
    mask = torch.full((SEQ, SEQ), float("-inf")).triu_(1)
    return x, freqs, mask


def main():
    block, args, transforms, linears = build_block()
    attn, ffn = block.attn, block.ffn
    rec = {f"arg_{k}": np.asarray(v) for k, v in ARGS.items()}
    for name, tr in transforms.items():
        rec[f"{name}.matrix_left"] = tr.matrix_left.data.numpy()
        rec[f"{name}.matrix_right"] = tr.matrix_right.data.numpy()
        rec[f"{name}.diag_scale"] = tr.diag_scale.data.numpy()
    with torch.no_grad():
        for name, (fl, tr) in linears.items():
            w = fl.apply_wclip(fl.apply_trans(fl.get_weight(), tr))      # _train_forward, deepseekv3_utils.py:77-84
            fl.weight_quantizer.find_params(w)
            wq = fl.weight_quantizer(w)
            scale = fl.weight_quantizer.scale.reshape(-1, 1)
            codes = torch.round(wq / scale).to(torch.int8)
            assert torch.equal(codes.float() * scale, wq)
            rec[f"{name}.w_packed"] = dq.pack_i4(codes).numpy()
            rec[f"{name}.s_w"] = scale.numpy()
            rec[f"{name}.clip_a"] = np.array([float(fl.act_quantizer.clip_factor_a_max),
                                              float(fl.act_quantizer.clip_factor_a_min)], np.float32)
    for name, t in (("attn_norm", block.attn_norm.weight), ("ffn_norm", block.ffn_norm.weight),
                    ("q_norm", attn.q_norm.weight), ("kv_norm", attn.kv_norm.weight), ("wkv_b", attn.wkv_b.weight),
                    ("gate", ffn.gate.weight)):
        rec[name] = t.data.numpy()
    rec["softmax_scale"] = np.float32(attn.softmax_scale)

    x, freqs, mask = fixture_input(args)
    caps = {}
    hooks = []

    def cap(mod, key, what="out"):
        if what == "out":
            hooks.append(mod.register_forward_hook(lambda m, i, o: caps.__setitem__(
                key, tuple(t.detach().clone() for t in o) if isinstance(o, tuple) else o.detach().clone())))
        else:
            # first call only: the transform module is called again on the WEIGHT inside FlatQuantizedLinear._train_forward
            hooks.append(mod.register_forward_hook(lambda m, i, o: (caps.setdefault(key, i[0].detach().clone()), None)[1]))  # a hook's return value replaces the output
    for nm in ("wq_a", "wkv_a", "wq_b", "wo"):
        cap(getattr(attn, nm), f"{nm}.out")
    cap(attn.wo_trans, "wo.in", "in")        # attention output before the online transform
    cap(attn, "attn.out")
    cap(ffn.gate, "gate.out")
    cap(ffn.shared_experts, "shared.out")
    cap(ffn, "ffn.out")
    with torch.no_grad():
        y = block(x, 0, freqs, mask)
    for h in hooks:
        h.remove()
    rec["x"], rec["y"] = x.numpy(), y.numpy()
    rec["freqs_real"], rec["freqs_imag"] = freqs.real.numpy(), freqs.imag.numpy()
    for k, v in caps.items():
        if isinstance(v, tuple):
            rec["gate.weights"], rec["gate.indices"] = v[0].numpy(), v[1].numpy()
        else:
            rec[k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "deepseek_block.npz"), **rec)
    print("block out absmax", float(y.abs().max()), "attn", float(caps["attn.out"].abs().max()),
          "ffn", float(caps["ffn.out"].abs().max()))
    print("expert load", np.bincount(rec["gate.indices"].reshape(-1), minlength=args.n_routed_experts))
    print("size", os.path.getsize(os.path.join(OUT, "deepseek_block.npz")))


if __name__ == "__main__":
    main()
