"""BASELINE.json configs[4] on the CPU: the host logic of the real-quant DeepSeek-V3 block (MLA + MoE) with the CUDA
entry points replaced by the oracle-backed stand-ins (tests/cpu_ops.py), against the reference's own fake-quant block
(tests/golden/make_golden_deepseek.py runs deepseek_v3/model.py + flatquant/model_tools/deepseekv3_utils.py)."""
import os
from types import SimpleNamespace

import numpy as np
import torch

import cpu_ops
from conftest import GOLDEN


def load_fixture():
    g = np.load(os.path.join(GOLDEN, "deepseek_block.npz"))
    args = SimpleNamespace(**{k[4:]: (g[k].item() if g[k].dtype.kind != "U" else str(g[k])) for k in g.files
                              if k.startswith("arg_")})
    params = {k: torch.from_numpy(g[k]) for k in g.files if not k.startswith("arg_")}
    return g, args, params


def _rel(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    d = np.abs(a - b)
    return float(d.max() / np.abs(b).max()), float(d.mean() / np.abs(b).mean())


def stage_errors(block, g, dev):
    """Every quantised stage on the golden input of that stage -> {stage: (max_rel, mean_rel)} vs the golden output."""
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    from flatquant_b200.deploy.transformers import modeling_llama as ML
    T = lambda a: torch.from_numpy(np.asarray(a)).to(dev)  # noqa: E731
    x = T(g["x"])
    b, s, _ = x.shape
    cos, sin = T(g["freqs_real"]), T(g["freqs_imag"])
    act = ML.ACT_DTYPE
    errs = {}
    with torch.no_grad():
        q1, kv = block.attn_in(MD._rms(x, block.w_attn_norm).to(act))
        errs["wq_a"], errs["wkv_a"] = _rel(q1.float().cpu(), g["wq_a.out"]), _rel(kv.float().cpu(), g["wkv_a.out"])
        errs["wq_b"] = _rel(block.attn_q(T(g["wq_a.out"]).to(act)).float().cpu(), g["wq_b.out"])
        core = block.attn_core(T(g["wq_b.out"]).to(act), T(g["wkv_a.out"]).to(act), 0, cos, sin)
        errs["attn_core(unquantised)"] = _rel(core.float().cpu(), g["wo.in"])
        errs["wo"] = _rel(block.attn_out(T(g["wo.in"]).to(act)).float().cpu(), g["wo.out"])
        h = x + T(g["attn.out"])
        xn = MD._rms(h, block.w_ffn_norm).to(act).reshape(b * s, -1)
        w, idx = block.gate(xn)
        errs["gate_weights"] = _rel(w.float().cpu(), g["gate.weights"])
        gi = T(g["gate.indices"])
        errs["gate_indices_mismatch"] = (float((idx.sort(-1)[0] != gi.sort(-1)[0]).float().mean()), 0.0)
        z = block.shared_expert(xn)
        errs["shared_expert(2 stages)"] = _rel(z.float().cpu().reshape(b, s, -1), g["shared.out"].reshape(b, s, -1))
        y = block.routed(xn, T(g["gate.weights"]).to(act), gi)
        errs["routed_experts(2 stages)"] = _rel(y.float().cpu().reshape(b, s, -1),
                                                g["ffn.out"].reshape(b, s, -1) - g["shared.out"].reshape(b, s, -1))
    return errs


def test_block_equals_reference_fake_quant_in_fp32(monkeypatch):
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    from flatquant_b200.deploy.transformers import modeling_llama as ML
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(cpu_ops, "EMULATE_FP16", False)
    monkeypatch.setattr(ML, "ACT_DTYPE", torch.float32)
    g, args, params = load_fixture()
    block = MD.FlatQuantDeepSeekBlock(args, params, "cpu")
    # our rotary table equals the reference's precompute_freqs_cis
    cos, sin = MD.precompute_freqs(args, g["x"].shape[1])
    assert np.allclose(cos.numpy(), g["freqs_real"], atol=1e-6) and np.allclose(sin.numpy(), g["freqs_imag"], atol=1e-6)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = block(x, 0, cos, sin)
    mx, mean = _rel(y.numpy(), g["y"])
    assert mx < 2e-5, (mx, mean)
    for k, (mx, mean) in stage_errors(block, g, "cpu").items():
        assert mx < 2e-5, (k, mx, mean)
    # wq_a / wkv_a own separate quantisers: two quantisations of one transformed input; w1 / w3 of the shared expert too
    assert len(block.g_qkv) == 2 and len(block.g_s13) == 2
    # decode: prefill s-1 tokens, then one token over the (16-bit, here fp32) latent cache == the prefill's last row
    s = x.shape[1]
    with torch.no_grad():
        block.kv_cache.zero_()
        block.pe_cache.zero_()
        block(x[:, :-1], 0, cos[:-1], sin[:-1])
        last = block(x[:, -1:], s - 1, cos[-1:], sin[-1:])
    assert _rel(last.numpy(), g["y"][:, -1:])[0] < 2e-5


# fp16 activations against the fp32 reference, one or two quantised stages at K <= 256: a single +-1 flip of a 4-bit code
# moves an output by several % of max|y| at these widths (cf. tests/test_gpu_parity.py::test_moe_routed_experts_configs4)
STAGE_TOL = {"default": (0.08, 0.03), "shared_expert(2 stages)": (0.25, 0.10), "routed_experts(2 stages)": (0.25, 0.10),
             "attn_core(unquantised)": (0.01, 0.005), "gate_weights": (0.01, 0.005), "gate_indices_mismatch": (0.05, 1.0)}


def check_stage_errors(errs):
    for k, (mx, mean) in errs.items():
        print(k, f"max {mx:.4f} mean {mean:.4f}")
    for k, (mx, mean) in errs.items():
        tmx, tmean = STAGE_TOL.get(k, STAGE_TOL["default"])
        assert mx < tmx and mean < tmean, (k, mx, mean)


def test_block_stages_teacher_forced_fp16(monkeypatch):
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    cpu_ops.install(monkeypatch)
    g, args, params = load_fixture()
    block = MD.FlatQuantDeepSeekBlock(args, params, "cpu")
    check_stage_errors(stage_errors(block, g, "cpu"))
    cos, sin = MD.precompute_freqs(args, g["x"].shape[1])
    with torch.no_grad():
        y = block(torch.from_numpy(g["x"]).half(), 0, cos, sin)
    c = torch.nn.functional.cosine_similarity(y.float().flatten(), torch.from_numpy(g["y"]).flatten(), dim=0)
    print("block output cosine", float(c))
    assert c > 0.97  # the residual stream dominates; the quantised branches are checked stage by stage above


def test_random_block_runs(monkeypatch):
    """The synthetic block bench.py --workload deepseek-block builds (here at reduced width) stays finite."""
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    cpu_ops.install(monkeypatch)
    args = SimpleNamespace(**dict(MD.V3_ARGS, dim=512, moe_inter_dim=256, n_heads=4, n_routed_experts=6,
                                  n_activated_experts=2, q_lora_rank=256, kv_lora_rank=64, qk_nope_head_dim=32,
                                  qk_rope_head_dim=32, v_head_dim=64, max_batch_size=1, max_seq_len=16))
    block = MD.FlatQuantDeepSeekBlock(args, MD.random_block_params(args, "cpu", seed=1), "cpu")
    cos, sin = MD.precompute_freqs(args, 10)
    with torch.no_grad():
        y = block(torch.randn(1, 10, 512).half(), 0, cos, sin)
    assert y.shape == (1, 10, 512) and torch.isfinite(y).all()
    assert len(block.g_qkv) == 1 and block.g_qkv[0].w.shape[0] == 256 + 64 + 32  # equal clips: wq_a and wkv_a fused


def test_grouped_routed_experts_equal_the_per_expert_loop(monkeypatch):
    """`FlatQuantRoutedExperts(grouped=True)` -- one grouped GEMM launch per stage, row offsets of the expert-sorted
    tokens kept on the device -- computes exactly what the per-expert loop computes (integer accumulation, same fp32
    epilogue), with and without expert parallelism, including experts that receive no token."""
    from flatquant_b200.moe import FlatQuantRoutedExperts
    cpu_ops.install(monkeypatch)
    torch.manual_seed(4)
    T_, dim, inter, E, top_k = 23, 256, 128, 6, 2
    mk = lambda n: (torch.eye(n) * 0.75 + torch.randn(n, n) * (0.5 / n ** 0.5)).half()  # noqa: E731
    f1, f2 = (mk(16), mk(16)), (mk(8), mk(16))
    w13 = torch.randint(0, 256, (E, 2 * inter, dim // 2), dtype=torch.uint8)
    w2 = torch.randint(0, 256, (E, dim, inter // 2), dtype=torch.uint8)
    s13, s2 = torch.rand(E, 2 * inter) * 0.02 + 0.01, torch.rand(E, dim) * 0.02 + 0.01
    d1, d2 = (torch.rand(dim) + 0.5).half(), (torch.rand(inter) + 0.5).half()
    x = torch.randn(T_, dim).half()
    probs = torch.tensor([0.4, 0.3, 0.0, 0.2, 0.1, 0.0]).repeat(T_, 1)  # experts 2 and 5 receive nothing
    indices = torch.multinomial(probs, top_k)
    weights = torch.rand(T_, top_k).half()
    for offset, sl in ((None, slice(0, E)), (0, slice(0, 3)), (3, slice(3, 6))):
        args = (w13[sl].contiguous(), s13[sl].contiguous(), w2[sl].contiguous(), s2[sl].contiguous(), f1, f2,
                (0.98, 0.97), (0.99, 0.98), d1, d2)
        loop = FlatQuantRoutedExperts(*args, expert_offset=offset, grouped=False)
        grouped = FlatQuantRoutedExperts(*args, expert_offset=offset, grouped=True)
        cpu_ops.CALLS.clear()
        y_loop = loop(x, weights, indices)
        n_loop = sum(op == "gemm_w4a8" for op, _ in cpu_ops.CALLS)
        cpu_ops.CALLS.clear()
        y_grouped = grouped(x, weights, indices)
        assert torch.equal(y_loop, y_grouped), offset
        launched = [op for op, _ in cpu_ops.CALLS if op.startswith("gemm")]
        assert launched == ["gemm_w4a8_grouped"] * 2 and n_loop > 2, (launched, n_loop)
    # no pair for the local experts: both paths return zeros without launching
    none = FlatQuantRoutedExperts(w13[2:3].contiguous(), s13[2:3].contiguous(), w2[2:3].contiguous(), s2[2:3].contiguous(),
                                  f1, f2, expert_offset=2, grouped=True)
    assert not none(x, weights, indices).any()
