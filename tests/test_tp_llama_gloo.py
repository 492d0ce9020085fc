"""Tensor-parallel Llama stack (BASELINE configs[3] semantics, vllm_custom/.../llama_flatquant.py:75-289) on CPU:
world_size 2 over gloo, the CUDA entry points replaced by the oracle-backed stand-ins (tests/cpu_ops.py).  Checks the
head / channel sharding, the per-rank row-parallel transforms, the per-rank KV cache and the two all-reduces per layer
against a lock-step serial restatement that sums the ranks' partial products by hand."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpu_ops

CFG = dict(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2,
           vocab_size=40, rope_theta=10000.0)
SEED, WORLD = 11, 2


def _install_stubs():
    from flatquant_b200 import ops
    for name in cpu_ops.FUNCTIONS:
        setattr(ops, name, getattr(cpu_ops, name))


def _ids():
    return torch.randint(0, CFG["vocab_size"], (2, 10), generator=torch.Generator().manual_seed(5))


def _worker(rank, world, port, ret):
    from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
    from flatquant_b200.deploy.transformers.modeling_llama import TensorParallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_stubs()
        model = FlatQuantLlamaForCausalLM.random_init(FlatQuantLlamaConfig(**CFG), torch.device("cpu"), seed=SEED,
                                                      tp=TensorParallel(world, rank))
        ids = _ids()
        with torch.no_grad():
            logits, cache = model(ids[:, :-1], max_length=ids.shape[1])
            step, _ = model(ids[:, -1:], past_key_values=cache)
        ret[rank] = (logits.float().numpy(), step.float().numpy(), tuple(cache.pages.shape))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tp2_stack_equals_lockstep_serial_restatement(monkeypatch):
    from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
    from flatquant_b200.deploy.transformers import modeling_llama as ML
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(WORLD, _free_port(), ret), nprocs=WORLD, join=True)
    # every rank ends with the same (all-reduced) logits; the per-rank cache holds that rank's kv heads only
    assert (ret[0][0] == ret[1][0]).all() and (ret[0][1] == ret[1][1]).all()
    assert ret[0][2][3] == CFG["num_key_value_heads"] // WORLD

    cpu_ops.install(monkeypatch)
    cfg = FlatQuantLlamaConfig(**CFG)
    models = []
    for r in range(WORLD):
        tp = ML.TensorParallel(WORLD, r)
        tp.all_reduce = lambda y: y  # partial products are summed by hand below
        tp.fused_reduction = False   # ... so the row-parallel linears must not reduce through a reducer object either
        models.append(FlatQuantLlamaForCausalLM.random_init(cfg, torch.device("cpu"), seed=SEED, tp=tp))
    m0 = models[0]
    # shapes of the shards (llama_flatquant.py:148-170, 106-112)
    at, mlp = m0.model.layers[0].self_attn, m0.model.layers[0].mlp
    assert at.num_heads == 2 and at.num_key_value_heads == 1 and at.q_proj.weight.shape == (256, 256)
    assert at.o_proj[1].weight.shape == (512, 128) and at.o_proj_trans.right_matrix.shape == (2, 2)
    assert mlp.gate_proj.weight.shape == (512, 256) and mlp.down_proj[0].left_matrix.shape[0] * \
        mlp.down_proj[0].right_matrix.shape[0] == 512
    # replicated parts agree across ranks, sharded parts differ
    a1 = models[1].model.layers[0].self_attn
    assert torch.equal(at.left_matrix, a1.left_matrix) and not torch.equal(at.q_proj.weight, a1.q_proj.weight)
    assert torch.equal(m0.lm_head.weight, models[1].lm_head.weight)

    ids = _ids()

    def run(tokens, caches, start):
        b, s = tokens.shape
        pos = torch.arange(start, start + s)[None].expand(b, s)
        cos, sin = ML.rope_cos_sin(ML.rope_inv_freq(cfg), pos, ML.ACT_DTYPE)
        h = m0.model.embed_tokens(tokens)
        for li in range(cfg.num_hidden_layers):
            layers = [m.model.layers[li] for m in models]
            x = layers[0].input_layernorm(h)
            h = h + sum(l.self_attn(x, cos, sin, c) for l, c in zip(layers, caches))
            x = layers[0].post_attention_layernorm(h)
            h = h + sum(l.mlp(x) for l in layers)
        return m0.lm_head(m0.model.norm(h))

    with torch.no_grad():
        caches = [m.build_cache(ids.shape[0], page_size=ids.shape[1], max_length=ids.shape[1]) for m in models]
        want = run(ids[:, :-1], caches, 0)
        want_step = run(ids[:, -1:], caches, ids.shape[1] - 1)
    assert (ret[0][0] == want.float().numpy()).all()       # fp16 a + b of two partials is order independent
    assert (ret[0][1] == want_step.float().numpy()).all()


def test_tp_placement_validation():
    from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
    from flatquant_b200.deploy.transformers.modeling_llama import TensorParallel, options_from_config
    import pytest
    cfg = FlatQuantLlamaConfig(**dict(CFG, num_hidden_layers=1))
    with pytest.raises(ValueError, match="num_attention_heads"):
        FlatQuantLlamaForCausalLM(options_from_config(cfg), cfg, TensorParallel(3, 0))
    m = FlatQuantLlamaForCausalLM(options_from_config(cfg), cfg, TensorParallel(4, 1))  # 2 kv heads on 4 ranks: replicated
    assert m.model.layers[0].self_attn.num_key_value_heads == 1 and m.model.layers[0].self_attn.num_heads == 1
