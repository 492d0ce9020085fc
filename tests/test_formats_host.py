"""the on-disk formats.  The fixture directory tests/golden/llama_ckpt was written by
the REFERENCE's save_flat_matrices / save_quantized_weights_with_safetensors (tests/golden/make_golden_llama.py); our
writers must reproduce those files tensor by tensor from the same parameters."""
import json
import os
import types

import pytest
import torch
from safetensors import safe_open
from safetensors.torch import load_file

from conftest import GOLDEN

CKPT = os.path.join(GOLDEN, "llama_ckpt")


class ParamTree(torch.nn.Module):
    """nn.Module tree whose named_parameters() are exactly `params` (dotted names -> tensors)."""

    def __init__(self, params):
        super().__init__()
        for name, t in params.items():
            mod = self
            *path, leaf = name.split(".")
            for part in path:
                if not hasattr(mod, part):
                    mod.add_module(part, torch.nn.Module())
                mod = getattr(mod, part)
            mod.register_parameter(leaf, torch.nn.Parameter(t.clone(), requires_grad=False))


def _reference_parameters():
    """The live parameters the reference exported: quantised weights as codes * scale (fp32), the rest as stored."""
    from flatquant_b200.deploy.functional.quantization import unpack_i4
    raw = load_file(os.path.join(CKPT, "model.safetensors"))
    params, quantizers = {}, {}
    for key, t in raw.items():
        if key.startswith("quantizer."):
            owner, kind = key[len("quantizer."):].rsplit(".", 1)
            setattr(quantizers.setdefault(owner, types.SimpleNamespace()), kind, t)
    for key, t in raw.items():
        if key.startswith("quantizer."):
            continue
        owner = key.rsplit(".", 1)[0]
        params[key] = unpack_i4(t).float() * quantizers[owner].scale if owner in quantizers else t.float()
    return raw, params, quantizers


def test_packed_int4_export_identical_to_reference(tmp_path):
    from flatquant_b200.flat_utils import save_quantized_weights_with_safetensors
    raw, params, quantizers = _reference_parameters()
    assert len(quantizers) == 14  # 7 linears x 2 layers
    args = types.SimpleNamespace(exp_dir=str(tmp_path), w_bits=4, model="tiny-llama-flatquant-fixture")
    save_quantized_weights_with_safetensors(args, ParamTree(params), quantizers)
    ours = load_file(os.path.join(tmp_path, "model.safetensors"))
    assert list(sorted(ours)) == list(sorted(raw))
    for k, t in raw.items():
        assert ours[k].dtype == t.dtype and ours[k].shape == t.shape and torch.equal(ours[k], t), k
    with safe_open(os.path.join(tmp_path, "model.safetensors"), "pt") as f, \
            safe_open(os.path.join(CKPT, "model.safetensors"), "pt") as r:
        assert f.metadata() == r.metadata()
    assert json.load(open(os.path.join(tmp_path, "quantization_config.json"))) == \
        json.load(open(os.path.join(CKPT, "quantization_config.json")))
    assert not os.path.exists(os.path.join(tmp_path, "model.safetensors.index.json"))  # one shard -> no index


def test_export_refuses_codes_outside_int4(tmp_path):
    """A weight that is not on its quantiser's grid range must not be silently wrapped into a nibble."""
    from flatquant_b200.flat_utils import quantized_state_dict
    q = types.SimpleNamespace(scale=torch.full((2, 1), 0.5), zero=torch.zeros(2, 1), maxq=torch.tensor(7))
    ok = quantized_state_dict(ParamTree({"l.weight": torch.tensor([[3.5, -4.0] * 8, [0.0, 1.0] * 8])}), {"l": q})
    assert ok["l.weight"].dtype == torch.uint8 and ok["l.weight"].shape == (2, 8)
    assert ok["l.weight"][0, 0] == (7 | (8 << 4))  # codes (7, -8): element 2j in the low nibble, two's complement
    # clamping is part of the format (flat_utils.py:127): 100.0 / 0.5 clamps to 7
    big = quantized_state_dict(ParamTree({"l.weight": torch.full((2, 16), 100.0)}), {"l": q})
    assert (big["l.weight"] == 0x77).all()


def test_flat_matrices_roundtrip_and_layout(tmp_path):
    """save_flat_matrices groups parameters in the reference's order (matrices, diagonal scales, weight clips,
    activation clips) under integer layer keys; load_flat_matrices restores them."""
    from flatquant_b200.flat_utils import load_flat_matrices, save_flat_matrices
    ref = torch.load(os.path.join(CKPT, "flat_matrices.pth"))

    class Block(ParamTree):
        calls = 0

        def rep_matrix_only(self):
            Block.calls += 1

    def make_model(fill):
        layers = []
        for i in sorted(ref):
            layer = torch.nn.Module()
            for part in ("self_attn", "mlp"):
                sub = {k[len(part) + 1:]: (v if fill else torch.zeros_like(v)) for k, v in ref[i].items()
                       if k.startswith(part + ".")}
                sub["unrelated.weight"] = torch.ones(2, 2)  # must not be saved
                layer.add_module(part, Block(sub))
            layers.append(layer)
        return types.SimpleNamespace(model=types.SimpleNamespace(layers=torch.nn.ModuleList(layers)))

    args = types.SimpleNamespace(exp_dir=str(tmp_path))
    save_flat_matrices(args, make_model(True))
    ours = torch.load(os.path.join(tmp_path, "flat_matrices.pth"))
    assert list(ours) == list(ref)
    for i in ref:
        assert list(ours[i]) == list(ref[i])  # same names in the same order
        assert all(torch.equal(ours[i][k], ref[i][k]) for k in ref[i])
    save_flat_matrices(args, make_model(True), rank=3)  # tensor-parallel shard naming
    assert os.path.exists(os.path.join(tmp_path, "flat_matrices_3.pth"))
    blank = make_model(False)
    load_flat_matrices(args, blank, path=CKPT)
    got = blank.model.layers[1].self_attn.get_parameter("ln_trans.matrix_left")
    assert torch.equal(got, ref[1]["self_attn.ln_trans.matrix_left"])
    assert Block.calls > 0


def test_reparameterize_ln_folds_diag_in_fp64():
    from flatquant_b200.flat_utils import reparameterize_ln
    torch.manual_seed(0)
    ln = types.SimpleNamespace(weight=torch.nn.Parameter(torch.rand(16).half()))
    trans = types.SimpleNamespace(diag_scale=torch.rand(16) + 0.5, use_diag=True)
    want = (ln.weight.data.double() * trans.diag_scale.double()).half()
    reparameterize_ln(ln, trans)
    assert torch.equal(ln.weight.data, want) and ln.weight.dtype == torch.float16 and trans.use_diag is False


def test_sharded_checkpoint_ingest(tmp_path):
    """from_pretrained reads the sharded layout (model-0000x-of-0000y.safetensors + index) as well."""
    from safetensors.torch import save_file

    from flatquant_b200.deploy.transformers import FlatQuantLlamaForCausalLM
    raw = load_file(os.path.join(CKPT, "model.safetensors"))
    keys = sorted(raw)
    half = len(keys) // 2
    names = ["model-00001-of-00002.safetensors", "model-00002-of-00002.safetensors"]
    save_file({k: raw[k] for k in keys[:half]}, os.path.join(tmp_path, names[0]))
    save_file({k: raw[k] for k in keys[half:]}, os.path.join(tmp_path, names[1]))
    with open(os.path.join(tmp_path, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": {k: names[i >= half] for i, k in enumerate(keys)}}, f)
    with open(os.path.join(tmp_path, "config.json"), "w") as f:
        f.write(open(os.path.join(CKPT, "config.json")).read())
    a = FlatQuantLlamaForCausalLM.from_pretrained(str(tmp_path)).state_dict()
    b = FlatQuantLlamaForCausalLM.from_pretrained(CKPT).state_dict()
    loaded = [k for k in a if "quantizer" not in k and not k.endswith(("o_proj.0.clip_factor_a_max",
              "o_proj.0.clip_factor_a_min", "down_proj.1.clip_factor_a_max", "down_proj.1.clip_factor_a_min"))]
    assert all(torch.equal(a[k], b[k]) for k in loaded if "diag_scale" not in k or "inp_trans_q" not in k
               and "inp_trans_v" not in k and "inp_trans_u" not in k)
    with pytest.raises(RuntimeError, match="not a local directory"):
        FlatQuantLlamaForCausalLM.from_pretrained("someone/some-hub-repo")
