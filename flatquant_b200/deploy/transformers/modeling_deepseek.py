"""Real-quant W4A4 DeepSeek-V3 block on the B200 kernels (BASELINE.json configs[4]: "DeepSeek-V3 single block incl. MoE
expert FlatLinears").  Semantics: deepseek_v3/model.py:393-733 (MLA with absorbed wkv_b, Gate, MoE, Block) wrapped by
flatquant/model_tools/deepseekv3_utils.py:139-470 (FlatQuantMLA, FlatQuantMoE, FlatQuantMoEExpert).

FlatLinears (W4A4, each with its own online Kronecker transform pair + per-channel `diag_scale`, and LAC clip factors):
    attention   wq_a, wkv_a   <- qkv_trans (one transform, two quantisations: the linears own separate quantisers)
                wq_b          <- wqb_trans (after q_norm)
                wo            <- wo_trans  (row-parallel in the reference)
                wkv_b stays unquantised (deepseekv3_utils.py:171-172) and is absorbed into q / the output as in model.py
    MoE         routed experts w1/w3 <- w1_trans with ONE shared quantiser, w2 <- routed_w2_trans with ONE shared
                quantiser (flatquant_b200/moe.py: transform + quantise once for all experts, [w1;w3] fused GEMM)
                shared expert  w1/w3 <- the same w1_trans, own quantisers; w2 <- w2_trans
The reference re-transforms, clips and RTN-quantises every weight on every forward (`_train_forward`, its wrappers
never switch to eval mode); here that is done once offline and the packed codes + scales are the block's parameters.
The latent KV cache is kept in 16 bit as in the reference (kv_cache / pe_cache, model.py:437-441).

Parameters arrive as a flat name -> tensor mapping (the layout of tests/golden/deepseek_block.npz, written by the
reference's own modules); RMSNorm, RoPE, the gate and the absorbed attention einsums are torch glue."""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import modeling_llama as ML
from ..nn import Linear4bit
from ...moe import FlatQuantRoutedExperts


def _act():
    return ML.ACT_DTYPE


def precompute_freqs(args, seqlen, device=None):
    """cos / sin of the rotary angles, [seqlen, rope_dim/2] (model.py:294-372; the YaRN correction only applies when
    max_seq_len > original_seq_len)."""
    dim = args.qk_rope_head_dim
    freqs = 1.0 / (args.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
    if args.max_seq_len > args.original_seq_len:
        def corr_dim(rot):
            return dim * math.log(args.original_seq_len / (rot * 2 * math.pi)) / (2 * math.log(args.rope_theta))
        lo = max(math.floor(corr_dim(args.beta_fast)), 0)
        hi = min(math.ceil(corr_dim(args.beta_slow)), dim - 1)
        ramp = torch.clamp((torch.arange(dim // 2, dtype=torch.float32, device=device) - lo) / max(hi - lo, 0.001), 0, 1)
        smooth = 1 - ramp
        freqs = freqs / args.rope_factor * (1 - smooth) + freqs * smooth
    ang = torch.outer(torch.arange(seqlen, dtype=torch.float32, device=device), freqs)
    return ang.cos(), ang.sin()


def apply_rotary(x, cos, sin):
    """Interleaved-pair rotation (model.py:375-390: view_as_complex over consecutive pairs).  x [b, s, h, d]."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xf[..., 0], xf[..., 1]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.stack((a * c - b * s, a * s + b * c), dim=-1).flatten(3).to(x.dtype)


def _rms(x, weight, eps=1e-6):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype)


class _Trans:
    """An eval-mode SVDDecomposeTransMatrix: x -> kron(x * diag_scale; left, right) (trans_utils.py:85-101)."""

    def __init__(self, params, name, dev):
        self.left = params[f"{name}.matrix_left"].to(dev, _act()).contiguous()
        self.right = params[f"{name}.matrix_right"].to(dev, _act()).contiguous()
        d = params.get(f"{name}.diag_scale")
        self.diag = None if d is None else d.to(dev, _act())
        self.n1, self.n2 = self.left.shape[0], self.right.shape[0]

    def scaled(self, x2):
        return x2 if self.diag is None else x2 * self.diag


def _linear(params, name, dev):
    w = params[f"{name}.w_packed"]
    lin = Linear4bit(w.shape[1] * 2, w.shape[0]).to(dev)
    lin.weight = w.to(dev).contiguous()
    lin.weight_scales = params[f"{name}.s_w"].to(dev, torch.float32).reshape(-1, 1)
    return lin


def _clip(params, name):
    c = params[f"{name}.clip_a"]
    return float(c[0]), float(c[1])


class FlatQuantDeepSeekBlock(torch.nn.Module):
    def __init__(self, args, params, device, tp=None):
        """`tp` (modeling_llama.TensorParallel): the reference's sharding of the block (deepseek_v3/model.py:205-264,
        404-441, 633-691) -- wq_b / wkv_b column-parallel over heads, wo row-parallel with this rank's own transform,
        routed experts range-partitioned, the shared expert column/row-parallel over its intermediate channels; `params`
        then holds this rank's shard (local shapes, local experts under their GLOBAL indices).  Two all-reduces per
        block: after wo, and one for routed + shared together (the reference issues them separately)."""
        super().__init__()
        self.args = a = args if not isinstance(args, dict) else SimpleNamespace(**args)
        self.tp = tp or ML.TensorParallel()
        dev = self.dev = torch.device(device)
        p = {k: (torch.as_tensor(v) if not isinstance(v, torch.Tensor) else v) for k, v in params.items()}
        self.n_heads = self.tp.shard(a.n_heads, "n_heads")  # local heads
        self.qk_head_dim = a.qk_nope_head_dim + a.qk_rope_head_dim
        self.softmax_scale = float(p["softmax_scale"]) if "softmax_scale" in p else self.qk_head_dim ** -0.5
        for n in ("attn_norm", "ffn_norm", "q_norm", "kv_norm", "gate"):
            setattr(self, f"w_{n}", p[n].to(dev, torch.float32))
        self.w_gate_bias = p["gate_bias"].to(dev, torch.float32) if "gate_bias" in p else None
        self.wkv_b = p["wkv_b"].to(dev, _act()).view(self.n_heads, a.qk_nope_head_dim + a.v_head_dim, a.kv_lora_rank)
        self.t_qkv, self.t_wqb, self.t_wo = (_Trans(p, n, dev) for n in ("qkv_trans", "wqb_trans", "wo_trans"))
        self.t_w1, self.t_w2, self.t_rw2 = (_Trans(p, n, dev) for n in ("w1_trans", "w2_trans", "routed_w2_trans"))
        grp = ML._group_by_clip
        self.g_qkv = grp(("wq_a", "wkv_a"), [_linear(p, n, dev) for n in ("wq_a", "wkv_a")],
                         [_clip(p, n) for n in ("wq_a", "wkv_a")])
        self.g_wqb = grp(("wq_b",), [_linear(p, "wq_b", dev)], [_clip(p, "wq_b")])
        self.g_wo = grp(("wo",), [_linear(p, "wo", dev)], [_clip(p, "wo")])
        self.g_s13 = grp(("w1", "w3"), [_linear(p, f"shared.{n}", dev) for n in ("w1", "w3")],
                         [_clip(p, f"shared.{n}") for n in ("w1", "w3")])
        self.g_s2 = grp(("w2",), [_linear(p, "shared.w2", dev)], [_clip(p, "shared.w2")])
        n_local = self.tp.shard(a.n_routed_experts, "n_routed_experts")
        first = self.tp.rank * n_local
        E = range(first, first + n_local)
        w13 = torch.stack([torch.cat([p[f"experts.{e}.w1.w_packed"], p[f"experts.{e}.w3.w_packed"]], 0) for e in E])
        s13 = torch.stack([torch.cat([p[f"experts.{e}.w1.s_w"], p[f"experts.{e}.w3.s_w"]], 0).reshape(-1) for e in E])
        w2 = torch.stack([p[f"experts.{e}.w2.w_packed"] for e in E])
        s2 = torch.stack([p[f"experts.{e}.w2.s_w"].reshape(-1) for e in E])
        c1, c2 = _clip(p, f"experts.{first}.w1"), _clip(p, f"experts.{first}.w2")  # one shared quantiser per stage (:413-415)
        for e in E:
            if _clip(p, f"experts.{e}.w1") != c1 or _clip(p, f"experts.{e}.w3") != c1 or _clip(p, f"experts.{e}.w2") != c2:
                raise RuntimeError("routed experts must share their activation quantisers (deepseekv3_utils.py:413-415)")
        sig = ML._sig
        self.routed = FlatQuantRoutedExperts(
            w13.to(dev).contiguous(), s13.to(dev, torch.float32).contiguous(), w2.to(dev).contiguous(),
            s2.to(dev, torch.float32).contiguous(), (self.t_w1.left, self.t_w1.right), (self.t_rw2.left, self.t_rw2.right),
            (sig(c1[0]), sig(c1[1])), (sig(c2[0]), sig(c2[1])), self.t_w1.diag, self.t_rw2.diag,
            expert_offset=first if self.tp.size > 1 else None)
        self.kv_cache = torch.zeros(a.max_batch_size, a.max_seq_len, a.kv_lora_rank, dtype=_act(), device=dev)
        self.pe_cache = torch.zeros(a.max_batch_size, a.max_seq_len, a.qk_rope_head_dim, dtype=_act(), device=dev)

    # ------------------------------------------------------------------------------------------------ FlatLinears
    @staticmethod
    def _fl(x, trans, groups):
        lead = x.shape[:-1]
        x2 = trans.scaled(x.reshape(-1, x.shape[-1]).to(_act())).contiguous()
        out = ML._flat_linears(x2, trans.left, trans.right, trans.n1, trans.n2, groups, True)
        return {k: v.reshape(*lead, -1) for k, v in out.items()}

    # ------------------------------------------------------------------------------------------------ attention
    def attn_in(self, xn):
        """wq_a and wkv_a on the normalised hidden state (one transform, one quantisation per quantiser)."""
        o = self._fl(xn, self.t_qkv, self.g_qkv)
        return o["wq_a"], o["wkv_a"]

    def attn_q(self, q1):
        return self._fl(_rms(q1, self.w_q_norm), self.t_wqb, self.g_wqb)["wq_b"]

    def attn_core(self, q, kv, start_pos, cos, sin, causal=True):
        """Absorbed MLA attention over the latent cache (deepseekv3_utils.py:182-231) -> [b, s, heads * v_head_dim]."""
        a = self.args
        b, s = q.shape[:2]
        end = start_pos + s
        q = q.view(b, s, self.n_heads, self.qk_head_dim)
        q_nope, q_pe = torch.split(q, [a.qk_nope_head_dim, a.qk_rope_head_dim], dim=-1)
        q_pe = apply_rotary(q_pe, cos, sin)
        kv, k_pe = torch.split(kv, [a.kv_lora_rank, a.qk_rope_head_dim], dim=-1)
        k_pe = apply_rotary(k_pe.unsqueeze(2), cos, sin).squeeze(2)
        q_nope = torch.einsum("bshd,hdc->bshc", q_nope, self.wkv_b[:, :a.qk_nope_head_dim])
        self.kv_cache[:b, start_pos:end] = _rms(kv, self.w_kv_norm)
        self.pe_cache[:b, start_pos:end] = k_pe
        kvc, pec = self.kv_cache[:b, :end], self.pe_cache[:b, :end]
        scores = (torch.einsum("bshc,btc->bsht", q_nope, kvc) + torch.einsum("bshr,btr->bsht", q_pe, pec)) * self.softmax_scale
        if causal and s > 1:
            mask = torch.full((s, end), float("-inf"), device=q.device).triu_(1 + start_pos)
            scores = scores + mask[None, :, None, :].to(scores.dtype)
        scores = scores.softmax(dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.einsum("bsht,btc->bshc", scores, kvc)
        return torch.einsum("bshc,hdc->bshd", o, self.wkv_b[:, -a.v_head_dim:]).flatten(2)

    def attn_out(self, o):
        return self.tp.all_reduce(self._fl(o, self.t_wo, self.g_wo)["wo"])

    def attention(self, xn, start_pos, cos, sin):
        q1, kv = self.attn_in(xn)
        return self.attn_out(self.attn_core(self.attn_q(q1), kv, start_pos, cos, sin))

    # ------------------------------------------------------------------------------------------------ MoE
    def gate(self, x2):
        """Top-k routing on the unquantised hidden state (model.py:563-595, softmax / sigmoid scoring, one group)."""
        a = self.args
        scores = F.linear(x2.float(), self.w_gate)
        scores = scores.softmax(dim=-1) if a.score_func == "softmax" else scores.sigmoid()
        select = scores if self.w_gate_bias is None else scores + self.w_gate_bias  # V3 (dim 7168): bias steers the choice only
        if getattr(a, "n_expert_groups", 1) > 1:
            raise NotImplementedError("grouped routing (n_expert_groups > 1)")
        indices = torch.topk(select, a.n_activated_experts, dim=-1)[1]
        weights = scores.gather(1, indices)
        if a.score_func == "sigmoid":
            weights = weights / weights.sum(dim=-1, keepdim=True)
        return (weights * a.route_scale).to(x2.dtype), indices

    def shared_expert(self, x2):
        gu = self._fl(x2, self.t_w1, self.g_s13)
        return self._fl(F.silu(gu["w1"]) * gu["w3"], self.t_w2, self.g_s2)["w2"]

    def moe(self, xn, routing=None):
        shape = xn.shape
        x2 = xn.reshape(-1, shape[-1]).to(_act())
        weights, indices = routing if routing is not None else self.gate(x2)
        y = self.routed(x2, weights, indices)
        return self.tp.all_reduce(y + self.shared_expert(x2)).view(shape)

    def forward(self, x, start_pos, cos, sin):
        x = x + self.attention(_rms(x, self.w_attn_norm), start_pos, cos, sin)
        return x + self.moe(_rms(x, self.w_ffn_norm))


# Public DeepSeek-V3 architecture (the reference's configs/config_671B.json is git-ignored)
V3_ARGS = dict(max_batch_size=1, max_seq_len=4096, dim=7168, moe_inter_dim=2048, n_heads=128, n_routed_experts=256,
               n_shared_experts=1, n_activated_experts=8, n_expert_groups=1, n_limited_groups=1, score_func="sigmoid",
               route_scale=2.5, q_lora_rank=1536, kv_lora_rank=512, qk_nope_head_dim=128, qk_rope_head_dim=64,
               v_head_dim=128, original_seq_len=4096, rope_theta=10000.0, rope_factor=40, beta_fast=32, beta_slow=1,
               mscale=1.0)


def random_block_params(args, device, seed=0, weight_scale=0.002, tp_size=1, tp_rank=0):
    """Synthetic parameters of one MoE block in the fixture layout (random INT4 codes, well-conditioned transform
    factors, unit diagonal scales, LAC clip factors at their initial 4.0) for benchmarks.  With tp_size > 1: this
    rank's shard -- replicated tensors are drawn from `seed`, sharded ones from a per-rank stream."""
    from ...function_utils import get_decompose_dim
    a = args if not isinstance(args, dict) else SimpleNamespace(**args)
    rep = torch.Generator(device=device).manual_seed(seed)
    own = torch.Generator(device=device).manual_seed(seed * 1009 + 7 + tp_rank)
    p = {}

    def trans(name, k, gen):
        n1, n2 = get_decompose_dim(k)
        for side, n in (("left", n1), ("right", n2)):
            p[f"{name}.matrix_{side}"] = (torch.eye(n, device=device) * 0.75 +
                                          torch.randn(n, n, device=device, generator=gen) * (0.5 / n ** 0.5))
        p[f"{name}.diag_scale"] = torch.ones(k, device=device)

    def linear(name, k, n, gen):
        p[f"{name}.w_packed"] = torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=device, generator=gen)
        p[f"{name}.s_w"] = torch.full((n, 1), weight_scale, device=device)
        p[f"{name}.clip_a"] = torch.tensor([4.0, 4.0])

    qk = a.qk_nope_head_dim + a.qk_rope_head_dim
    heads = a.n_heads // tp_size
    shared_inter = a.n_shared_experts * a.moe_inter_dim // tp_size
    n_local = a.n_routed_experts // tp_size
    # replicated: the input-side transforms, wq_a / wkv_a (plain Linear in model.py:418-421), norms, the gate
    trans("qkv_trans", a.dim, rep), trans("wqb_trans", a.q_lora_rank, rep), trans("w1_trans", a.dim, rep)
    trans("routed_w2_trans", a.moe_inter_dim, rep)
    linear("wq_a", a.dim, a.q_lora_rank, rep), linear("wkv_a", a.dim, a.kv_lora_rank + a.qk_rope_head_dim, rep)
    p["gate"] = torch.randn(a.n_routed_experts, a.dim, device=device, generator=rep) * 0.05
    for n, k in (("attn_norm", a.dim), ("ffn_norm", a.dim), ("q_norm", a.q_lora_rank), ("kv_norm", a.kv_lora_rank)):
        p[n] = torch.ones(k, device=device)
    # sharded: heads, the row-parallel transforms (sized by the local K), the shared expert's channels, the experts
    trans("wo_trans", heads * a.v_head_dim, own), trans("w2_trans", shared_inter, own)
    linear("wq_b", a.q_lora_rank, heads * qk, own), linear("wo", heads * a.v_head_dim, a.dim, own)
    linear("shared.w1", a.dim, shared_inter, own), linear("shared.w3", a.dim, shared_inter, own)
    linear("shared.w2", shared_inter, a.dim, own)
    for e in range(tp_rank * n_local, (tp_rank + 1) * n_local):
        linear(f"experts.{e}.w1", a.dim, a.moe_inter_dim, own), linear(f"experts.{e}.w3", a.dim, a.moe_inter_dim, own)
        linear(f"experts.{e}.w2", a.moe_inter_dim, a.dim, own)
    p["wkv_b"] = torch.randn(heads * (a.qk_nope_head_dim + a.v_head_dim), a.kv_lora_rank, device=device,
                             generator=own) * a.kv_lora_rank ** -0.5
    return p
