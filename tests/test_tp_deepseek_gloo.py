"""Tensor / expert parallel DeepSeek-V3 block on CPU: world_size 2 over gloo with the oracle-backed stand-ins
(tests/cpu_ops.py).  Heads, the shared expert's channels and the routed experts are split as in the reference
(deepseek_v3/model.py:205-264, 404-441, 657-660); the result must equal a lock-step serial restatement that sums the
ranks' partial products by hand."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpu_ops

SEED, WORLD = 5, 2


def _args():
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    return SimpleNamespace(**dict(MD.V3_ARGS, dim=512, moe_inter_dim=256, n_heads=4, n_routed_experts=6,
                                  n_activated_experts=2, q_lora_rank=256, kv_lora_rank=64, qk_nope_head_dim=32,
                                  qk_rope_head_dim=32, v_head_dim=64, max_batch_size=2, max_seq_len=16,
                                  score_func="softmax", route_scale=1.0))


def _x():
    return torch.randn(2, 7, 512, generator=torch.Generator().manual_seed(3)).half()


def _block(rank, tp):
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    a = _args()
    return MD.FlatQuantDeepSeekBlock(a, MD.random_block_params(a, "cpu", seed=SEED, tp_size=WORLD, tp_rank=rank), "cpu", tp)


def _worker(rank, world, port, ret, grouped):
    from flatquant_b200 import ops
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    from flatquant_b200.deploy.transformers.modeling_llama import TensorParallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    # grouped (the default): one launch per expert stage, local experts only; "0": the per-expert loop
    os.environ["FLATQ_GROUPED_MOE"] = "1" if grouped else "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name in cpu_ops.FUNCTIONS:
            setattr(ops, name, getattr(cpu_ops, name))
        block = _block(rank, TensorParallel(world, rank))
        assert block.routed.grouped == bool(grouped)
        x = _x()
        cos, sin = MD.precompute_freqs(_args(), x.shape[1])
        with torch.no_grad():
            y = block(x[:, :-1], 0, cos[:-1], sin[:-1])
            last = block(x[:, -1:], x.shape[1] - 1, cos[-1:], sin[-1:])
        ret[rank] = (y.float().numpy(), last.float().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("grouped", [False, True])
def test_tp2_block_equals_lockstep_serial_restatement(monkeypatch, grouped):
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    from flatquant_b200.deploy.transformers.modeling_llama import TensorParallel
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(WORLD, _free_port(), ret, grouped), nprocs=WORLD, join=True)
    assert (ret[0][0] == ret[1][0]).all() and (ret[0][1] == ret[1][1]).all()

    cpu_ops.install(monkeypatch)
    blocks = []
    for r in range(WORLD):
        tp = TensorParallel(WORLD, r)
        tp.all_reduce = lambda y: y
        blocks.append(_block(r, tp))
    b0 = blocks[0]
    assert b0.n_heads == 2 and b0.routed.n_experts == 3 and blocks[1].routed.expert_offset == 3
    assert b0.g_wo[0].w.shape == (512, 64) and b0.t_wo.n1 * b0.t_wo.n2 == 128  # row-parallel: transform of the local K
    assert b0.g_s2[0].w.shape == (512, 64)                                      # shared expert: 256 / 2 channels
    x = _x()
    cos, sin = MD.precompute_freqs(_args(), x.shape[1])

    def run(xs, start, c, s):
        xn = MD._rms(xs, b0.w_attn_norm)
        h = xs + sum(b.attention(xn, start, c, s) for b in blocks)
        xn = MD._rms(h, b0.w_ffn_norm)
        return h + sum(b.moe(xn) for b in blocks)
    with torch.no_grad():
        want = run(x[:, :-1], 0, cos[:-1], sin[:-1])
        want_last = run(x[:, -1:], x.shape[1] - 1, cos[-1:], sin[-1:])
    assert (ret[0][0] == want.float().numpy()).all()
    assert (ret[0][1] == want_last.float().numpy()).all()
    # every (token, slot) pair is served by exactly one rank
    xn = MD._rms(x.reshape(-1, 512), b0.w_ffn_norm)
    _, idx = b0.gate(xn)
    assert ((idx < 3) | (idx >= 3)).all() and int((idx < 3).sum()) + int((idx >= 3).sum()) == idx.numel()
