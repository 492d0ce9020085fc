"""kronecker_matmul -- the online transform (mirror of flatquant/flat_utils.py:6-17)."""
import torch

from . import ops

_KERNEL_MAX = 256
_F16_MAX = 65504.0


def _shape_ok(n1, n2):
    return (n2 % 16 == 0 and n1 <= _KERNEL_MAX and n2 <= _KERNEL_MAX
            and (n1 + 15) // 16 * 32 <= (512 if n2 <= 128 else 256))


def _inference_act(x):
    """CUDA, 16-bit, no autograd: an inference activation -- the sm_100a kernel's job, never a silent cuBLAS call."""
    return x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and not torch.is_grad_enabled()


def as_kernel_f16(x):
    """fp16 view of a 16-bit activation for the kernels.  bf16 -> fp16 is exact (8 -> 11 mantissa bits) for finite values
    up to 65504; larger magnitudes saturate there instead of becoming inf."""
    if x.dtype == torch.float16:
        return x.contiguous()
    return x.clamp(-_F16_MAX, _F16_MAX).to(torch.float16).contiguous()


def kronecker_matmul(x, hadL, hadR):
    """x @ kron(hadL, hadR), evaluated per token as hadL^T . reshape(x, n1, n2) . hadR.

    Inference activations (CUDA, fp16 or bf16, no grad) run on the sm_100a kernel (flatq_kron_transform; bf16 is
    converted exactly to fp16 on the way in, the factors are rounded to the activation dtype like the reference's
    `.to(inp)`, the result is rounded back to bf16) and RAISE when the factor sizes have no kernel -- they are never
    routed to cuBLAS behind the caller's back.  What remains on torch.matmul, as in the reference, is not the inference
    path: the offline fp64 weight reparameterisation, calibration-time autograd, fp32 evaluation and CPU tensors."""
    n1, n2 = hadL.shape[0], hadR.shape[0]
    if _inference_act(x):
        if not _shape_ok(n1, n2):
            raise RuntimeError(f"kronecker_matmul: no sm_100a kernel for factor sizes ({n1}, {n2}) (need n2 % 16 == 0, "
                               "n1, n2 <= 256); fp32 activations take the reference's torch path")
        xc = as_kernel_f16(x)
        z = ops.kron_transform(xc, hadL.to(x).to(xc).contiguous(), hadR.to(x).to(xc).contiguous(), n1, n2)
        return z.view(x.shape).to(x.dtype)
    lead = x.shape
    t = x.reshape(-1, n1, n2)
    return (hadL.transpose(0, 1) @ (t @ hadR)).reshape(lead)


# ------------------------------------------------------------------------------------------------------------------
# On-disk formats (flatquant/flat_utils.py:20-204).  Offline host code, torch only: these
# run once per checkpoint, on whatever device the calibrated model lives on; nothing here is on the inference path.
# The readers of the packed export are flatquant_b200.deploy.transformers.FlatQuantLlamaForCausalLM.from_pretrained
# and the reference's own loader (deploy/transformers/modeling_llama.py:356-542): file names, tensor names, dtypes and
# metadata are the contract and are checked tensor by tensor against files written by the reference
# (tests/test_formats_host.py).
# ------------------------------------------------------------------------------------------------------------------
FLAT_MATRIX_NAMES = ("trans.matrix", "trans.diag_scale", "clip_factor_w", "clip_factor_a")


def reparameterize_ln(ln, trans):
    """Fold a transform's per-channel `diag_scale` into the preceding RMSNorm weight in fp64 and switch the
    transform's diagonal off (flat_utils.py:20-27)."""
    w = ln.weight.data
    ln.weight.data = (w.double() * trans.diag_scale.double()).to(w.dtype)
    trans.use_diag = False


def reparameterize_model(model):
    """Bake every layer's transforms into its weights and fold the diagonal scales into the layer norms
    (flat_utils.py:30-41)."""
    for layer in model.model.layers[:model.config.num_hidden_layers]:
        layer.self_attn.reparameterize()
        layer.mlp.reparameterize()
        for ln, trans in ((layer.input_layernorm, layer.self_attn.ln_trans),
                          (layer.post_attention_layernorm, layer.mlp.up_gate_trans)):
            if trans is not None and trans.add_diag:
                reparameterize_ln(ln, trans)
    return model


def _exp_path(args, path, name):
    import os
    return os.path.join(args.exp_dir if path is None else path, name)


def save_parametrized_checkpoint(model, args):
    """{layer index: layer.state_dict()} -> <exp_dir>/parametrized_paras.pth (flat_utils.py:44-50)."""
    torch.save({i: layer.state_dict() for i, layer in enumerate(model.model.layers)},
               _exp_path(args, None, "parametrized_paras.pth"))


def load_flat_parameters(args, model, path=None):
    """Resume calibration state from flat_parameters.pth (flat_utils.py:53-62)."""
    saved = torch.load(_exp_path(args, path, "flat_parameters.pth"))
    for i in range(len(saved)):
        model.model.layers[i].load_state_dict(saved[i], strict=False)
    return model


def _matrices_to_eval(layer):
    layer.self_attn.rep_matrix_only()
    layer.mlp.rep_matrix_only()


def save_flat_matrices(args, model, rank=None):
    """Eval-mode transform matrices, diagonal scales and clip factors of every layer ->
    <exp_dir>/flat_matrices.pth, or flat_matrices_<rank>.pth for a tensor-parallel shard (flat_utils.py:65-79)."""
    from .function_utils import get_paras_dict_by_name
    out = {}
    for i, layer in enumerate(model.model.layers):
        _matrices_to_eval(layer)
        out[i] = get_paras_dict_by_name(layer, required_names=list(FLAT_MATRIX_NAMES))
    torch.save(out, _exp_path(args, None, "flat_matrices.pth" if rank is None else f"flat_matrices_{rank}.pth"))


def load_flat_matrices(args, model, path=None):
    """Inverse of save_flat_matrices: layers are switched to eval-mode matrices first so the keys exist
    (flat_utils.py:82-94)."""
    saved = torch.load(_exp_path(args, path, "flat_matrices.pth"))
    for i in range(len(saved)):
        _matrices_to_eval(model.model.layers[i])
        model.model.layers[i].load_state_dict(saved[i], strict=False)
    return model


def quantized_state_dict(model, quantizers, sym=True):
    """The tensors of the packed-INT4 export (flat_utils.py:107-146): weights of quantised linears as 4-bit codes
    packed two per byte (deploy.functional.pack_i4), every other parameter in fp16, and per quantised linear
    `quantizer.<name>.{scale,zero,maxq}`."""
    from .deploy.functional.quantization import pack_i4
    sd = {}
    for name, param in model.named_parameters():
        owner = name.rsplit(".", 1)[0] if name.endswith((".weight", ".bias")) else name
        if owner in quantizers and "weight" in name:
            q = quantizers[owner]
            scale, zero, maxq = (t.to(param.device) for t in (q.scale, q.zero, q.maxq))
            codes = torch.round(param / scale)
            codes = torch.clamp(codes, -(maxq + 1), maxq) if sym else torch.clamp(codes + zero, 0, maxq)
            sd[name] = pack_i4(codes.to(torch.int8)).contiguous()
        else:
            sd[name] = param.to(torch.float16).contiguous()
    for owner, q in quantizers.items():
        sd[f"quantizer.{owner}.scale"] = q.scale.contiguous()
        for extra in ("zero", "maxq"):
            t = getattr(q, extra, None)
            if t is not None:
                sd[f"quantizer.{owner}.{extra}"] = t.contiguous()
    return sd


def save_quantized_weights_with_safetensors(args, model, quantizers, sym=True):
    """Packed-INT4 safetensors export (flat_utils.py:97-204): `model*.safetensors` shards of at most 5 GB (+
    `model.safetensors.index.json` when sharded), the quantisation summary as metadata of the first shard and as
    `quantization_config.json`."""
    import json
    import os

    from huggingface_hub import split_torch_state_dict_into_shards
    from safetensors.torch import save_file
    sd = quantized_state_dict(model, quantizers, sym)
    plan = split_torch_state_dict_into_shards(sd, max_shard_size="5GB", filename_pattern="model{suffix}.safetensors")
    os.makedirs(args.exp_dir, exist_ok=True)
    summary = {"w_bits": args.w_bits, "model_name": args.model, "symmetric": sym, "format": "packed_int4"}
    for n, (fname, names) in enumerate(plan.filename_to_tensors.items()):
        save_file({k: sd[k] for k in names}, os.path.join(args.exp_dir, fname),
                  metadata={"quantization_config": json.dumps(summary)} if n == 0 else None)
    if plan.is_sharded:
        with open(os.path.join(args.exp_dir, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": getattr(plan, "metadata", {}), "weight_map": plan.tensor_to_filename}, f, indent=2)
    with open(os.path.join(args.exp_dir, "quantization_config.json"), "w") as f:
        json.dump(dict(summary, sharded=plan.is_sharded), f, indent=2)
