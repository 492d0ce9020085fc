"""Which kernel configurations do the model stacks (SURVEY.md section 8 f-1, configs[2] / configs[4] fixtures) drive,
and has each one been through a GPU parity test?  The CPU stand-ins (tests/cpu_ops.py, STRICT) apply the C ABI's own
argument checks and record (op, configuration) per call; this test pins that list against the table below, so a
change in the host code that reaches a kernel configuration no `-m gpu` test covers fails HERE, before the GPU run."""
import os

import numpy as np
import torch

import cpu_ops
from conftest import GOLDEN

# kernel class -> the `-m gpu` test that runs it on the B200 against the oracle
GPU_VERIFIED = {
    "hmma two-factor": "test_gpu_parity.py::test_transform_quant_codes[(32,48),(128,224),(8,16),(14,16)] (padded n1 included)",
    "hmma left-only": "test_gpu_parity.py::test_kernel_level_api (block_matmul; head counts other than 16 / 32 / 64)",
    "hmma right-only": "test_zz_llama_model_gpu.py::test_right_only_transform_and_matmul_kv_cache (hardware-verified round 2; the K / Q cache transform)",
    "tcgen05": "test_gpu_parity.py::test_transform_quant_codes[(64,64),(112,128),(86,128),(56,64),(32,64),(64,128),(128,128),(48,64)], "
               "test_left_only_transform_on_the_tcgen05_kernel (o_proj over 16 / 32 / 64 heads x 128)",
    "rowquant": "test_gpu_parity.py::test_quantize_rows_deploy_semantics, test_empty_and_maximum_shapes",
    "gemv": "test_gpu_parity.py::test_decode_split_k_path[m=1,2]",
    "skinny": "test_gpu_parity.py::test_decode_gemv_and_skinny_gemm_bit_exact[m=3..32], test_decode_split_k_path[m=5..17]",
    "split-k": "test_gpu_parity.py::test_decode_split_k_path[m=40,64]",
    "fused-decode": "test_gpu_parity.py::test_one_launch_decode_flatlinear (transform + quantise in the GEMV's prologue, 1-2 tokens)",
    "1sm-128": "test_gpu_parity.py::test_gemm_i4_exact[(100,200,192),(128,256,4096)], __graft_entry__.smoke (96 tokens)",
    "2sm": "test_gpu_parity.py::test_prefill_gemm_variants_bit_identical, __graft_entry__.smoke (320 tokens)",
    "kv_quant": "test_gpu_parity.py::test_kv_quant_golden_bit_exact, test_paged_kv_append_and_decode",
    "kv_decode_i4": "test_gpu_parity.py::test_paged_kv_append_and_decode, test_kv_cache_module_prefill_then_decode",
    "kv_decode_fused": "test_zzz_kv_fused_gpu.py::test_decode_step_with_the_kv_front_folded_into_the_attention_kernel",
}


def _classify(op, cfg):
    if op in ("gemm_w4a8", "gemm_w4a8_silu"):
        return cfg[0]
    if op == "linear_fused_decode":
        return "fused-decode"
    if op in ("kv_quant", "kv_decode_i4", "kv_decode_fused"):
        return op
    if op == "act_quant" and cfg[0] == "rowquant":
        return "rowquant"
    if op == "act_quant":
        kernel, _, _, has_l, has_r = cfg
    else:
        kernel, (_, _, has_l, has_r) = "hmma", cfg
    if kernel == "tcgen05":
        return "tcgen05"
    return "hmma two-factor" if has_l and has_r else "hmma left-only" if has_l else "hmma right-only"


def _classes(calls):
    return {_classify(op, cfg) for op, cfg in calls}


def test_llama_stack_kernel_configurations(monkeypatch):
    from flatquant_b200.deploy.transformers import FlatQuantLlamaForCausalLM
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(cpu_ops, "CALLS", [])
    g = np.load(os.path.join(GOLDEN, "llama_model.npz"))
    model = FlatQuantLlamaForCausalLM.from_pretrained(os.path.join(GOLDEN, "llama_ckpt"))
    ids = torch.from_numpy(g["ids"])
    with torch.no_grad():
        _, cache = model(ids[:, :-1], max_length=ids.shape[1])
        model(ids[:, -1:], past_key_values=cache)
    used = _classes(cpu_ops.CALLS)
    # prefill 30 tokens -> weight-streaming mma.sync GEMM (gemm_w4a8_skinny.cuh), decode 2 tokens -> GEMV; (16,32) / (32,32) / padded (4,128)+identity
    # Kronecker pairs on the HMMA kernel; K.T_k and Q.T_k^-T as right-only 128x128 transforms; paged INT4 KV
    assert used == {"hmma two-factor", "hmma right-only", "skinny", "gemv", "kv_quant", "kv_decode_i4", "kv_decode_fused"}, used
    assert used <= set(GPU_VERIFIED)
    configs = {cfg for op, cfg in cpu_ops.CALLS if op == "act_quant"}
    assert configs == {("hmma", 16, 32, True, True), ("hmma", 4, 128, True, True), ("hmma", 32, 32, True, True)}, configs


def test_llama8b_shape_kernel_configurations(monkeypatch):
    """The Llama-3-8B architecture (2 layers here) as bench.py --workload llama8b-model builds it."""
    from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(cpu_ops, "CALLS", [])
    cfg = FlatQuantLlamaConfig.llama3_8b(num_hidden_layers=1, vocab_size=32)
    model = FlatQuantLlamaForCausalLM.random_init(cfg, torch.device("cpu"), seed=0)
    ids = torch.randint(0, 32, (1, 6), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        _, cache = model(ids, max_length=8)
        model(ids[:, -1:], past_key_values=cache)
    configs = {cfg for op, cfg in cpu_ops.CALLS if op == "act_quant"}
    # hidden 4096 -> (64,64) and inter 14336 -> (112,128) on the tcgen05 kernel; o_proj = left-only 32 heads x 128, also
    # on the tcgen05 kernel (identity right factor) since round 2
    assert configs == {("tcgen05", 64, 64, True, True), ("tcgen05", 112, 128, True, True), ("tcgen05", 32, 128, True, False)}, configs
    assert _classes(cpu_ops.CALLS) <= set(GPU_VERIFIED)


def test_deepseek_block_kernel_configurations(monkeypatch):
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    from test_deepseek_block_host import load_fixture
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(cpu_ops, "CALLS", [])
    g, args, params = load_fixture()
    block = MD.FlatQuantDeepSeekBlock(args, params, "cpu")
    x = torch.from_numpy(g["x"]).half()
    cos, sin = MD.precompute_freqs(args, x.shape[1])
    with torch.no_grad():
        block(x[:, :-1], 0, cos[:-1], sin[:-1])
        block(x[:, -1:], x.shape[1] - 1, cos[-1:], sin[-1:])
    used = _classes(cpu_ops.CALLS)
    assert used == {"hmma two-factor", "skinny", "gemv"}, used
    configs = {cfg[1:3] for op, cfg in cpu_ops.CALLS if op == "act_quant"}
    assert configs == {(16, 16), (8, 16)}, configs  # dim 256; q_lora / wo / moe_inter 128
