"""Routed-expert FlatLinears of the DeepSeek-V3 block (BASELINE configs[4]); semantics of
flatquant/model_tools/deepseekv3_utils.py:365-455 (`FlatQuantMoEExpert`, `FlatQuantMoE.forward`):

    x_t = w1_trans(x)                                   # ONE transform for all experts (:432-433)
    for each expert e with tokens idx_e:
        h = silu(w1_e(x_t[idx_e])) * w3_e(x_t[idx_e])   # w1/w3 share the activation quantiser (:377-378, :424)
        y[idx_e] += w2_e(routed_w2_trans(h)) * gate_weight

B200 arrangement with the existing kernels: the transform AND the per-token quantisation run once for all tokens
(`act_quant`), the experts gather rows of the int8 operand (a byte gather, not a re-quantisation), w1 and w3 are one
fused GEMM per expert ([2*inter, K] packed), the second transform+quantise runs on the gathered rows of all experts
at once (per-token scales make it order independent), then one GEMM per expert.  Small per-expert token counts take
the split-K route automatically.

`grouped=True` (the default; FLATQ_GROUPED_MOE=0 or grouped=False selects the per-expert loop) replaces the per-expert GEMM loops by ONE grouped launch per stage
(`ops.gemm_w4a8_grouped`: the (expert, tile) schedule is built on the device from the row offsets of the expert-sorted
tokens, csrc/grouped_sched.h), which also removes the host read of the routing result (`bincount().tolist()`), so a
step no longer synchronises with the device.  Integer accumulation + the same fp32 epilogue: results are bit-identical
to the loop.  Hardware-verified in round 2 (tests/test_zzz_grouped_gemm_gpu.py,
profiles/r02_first_gpu_call.md): the DeepSeek-V3 block step went from 34.1 ms (loop) to 12.8 ms (grouped).
"""
from __future__ import annotations

import os

import torch

from . import ops


class FlatQuantRoutedExperts:
    def __init__(self, w13_packed, s_w13, w2_packed, s_w2, w1_factors, w2_factors, clip1=(1.0, 1.0), clip2=(1.0, 1.0),
                 diag1=None, diag2=None, expert_offset=None, grouped=None):
        """w13_packed [E, 2*inter, dim/2] u8 (rows: w1 then w3), s_w13 [E, 2*inter] f32, w2_packed [E, dim, inter/2],
        s_w2 [E, dim]; w1_factors / w2_factors = (left, right) fp16 Kronecker pairs; clip = post-sigmoid factors;
        diag1 [dim] / diag2 [inter]: the transforms' per-channel `diag_scale` (add_diag; the DeepSeek wrappers keep it
        online, trans_utils.py:86-90), applied to the activation in front of the Kronecker factors;
        expert_offset: expert parallelism -- this rank holds global experts [offset, offset + E) and skips the
        (token, slot) pairs routed elsewhere (deepseek_v3/model.py:657-660, 677-683);
        grouped: one grouped GEMM launch per stage instead of one GEMM per expert (None: FLATQ_GROUPED_MOE env)."""
        self.w13, self.s13, self.w2, self.s2 = w13_packed, s_w13, w2_packed, s_w2
        self.f1, self.f2, self.clip1, self.clip2 = w1_factors, w2_factors, clip1, clip2
        self.diag1, self.diag2, self.expert_offset = diag1, diag2, expert_offset
        self.n_experts = w13_packed.shape[0]
        self.inter = w2_packed.shape[2] * 2
        self.dim = w13_packed.shape[2] * 2
        self.grouped = (os.environ.get("FLATQ_GROUPED_MOE", "1") != "0") if grouped is None else bool(grouped)

    def __call__(self, x: torch.Tensor, weights: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        """x [T, dim] fp16, weights / indices [T, top_k] from the (unquantised) gate -> y [T, dim] fp16."""
        T, top_k = indices.shape
        L1, R1 = self.f1
        dt = x.dtype
        if self.diag1 is not None:
            x = x * self.diag1
        xq8, sx = ops.act_quant(x.contiguous(), L1, R1, L1.shape[0], R1.shape[0], *self.clip1)  # once for all experts
        # sort the (token, slot) pairs by expert
        flat_e = indices.reshape(-1)
        if self.expert_offset is None:
            order = torch.argsort(flat_e, stable=True)
        else:  # keep the pairs of the local experts only
            flat_e = flat_e - self.expert_offset
            mine = ((flat_e >= 0) & (flat_e < self.n_experts)).nonzero().squeeze(1)
            order = mine[torch.argsort(flat_e[mine], stable=True)]
            flat_e = flat_e[mine]
        tok = (order // top_k)
        if self.grouped:
            return self._grouped(x, xq8, sx, weights, order, tok, flat_e, T, dt)
        counts = torch.bincount(flat_e, minlength=self.n_experts).tolist()
        xs8, sxs = xq8[tok].contiguous(), sx[tok].contiguous()
        h = torch.empty((tok.numel(), self.inter), dtype=dt, device=x.device)
        off = 0
        for e, c in enumerate(counts):
            if c:
                gu = ops.gemm_w4a8(xs8[off:off + c], sxs[off:off + c], self.w13[e], self.s13[e], out_dtype=dt)
                g, u = gu[:, :self.inter], gu[:, self.inter:]
                h[off:off + c] = torch.nn.functional.silu(g.float()).mul_(u.float()).to(dt)
            off += c
        if tok.numel() == 0:  # no token of this batch is routed to a local expert
            return torch.zeros((T, self.dim), dtype=dt, device=x.device)
        L2, R2 = self.f2
        if self.diag2 is not None:
            h = h * self.diag2
        hq8, sh = ops.act_quant(h, L2, R2, L2.shape[0], R2.shape[0], *self.clip2)  # all experts' rows in one launch
        out = torch.empty((tok.numel(), self.dim), dtype=dt, device=x.device)
        off = 0
        for e, c in enumerate(counts):
            if c:
                ops.gemm_w4a8(hq8[off:off + c], sh[off:off + c], self.w2[e], self.s2[e], out_dtype=dt, out=out[off:off + c])
            off += c
        w_sorted = weights.reshape(-1)[order].to(torch.float32)
        y = torch.zeros((T, self.dim), dtype=torch.float32, device=x.device)
        y.index_add_(0, tok, out.float() * w_sorted[:, None])
        return y.to(dt)

    def _grouped(self, x, xq8, sx, weights, order, tok, local_e, T, dt):
        """Both stages as one grouped launch each; nothing below reads device data on the host."""
        n_pairs = tok.numel()
        if n_pairs == 0:
            return torch.zeros((T, self.dim), dtype=dt, device=x.device)
        counts = torch.bincount(local_e, minlength=self.n_experts)
        offsets = torch.zeros(self.n_experts + 1, dtype=torch.int32, device=x.device)
        offsets[1:] = torch.cumsum(counts, 0)
        hint = -(-n_pairs // self.n_experts)  # mean rows per expert (from shapes only)
        xs8, sxs = xq8[tok].contiguous(), sx[tok].contiguous()
        gu = ops.gemm_w4a8_grouped(xs8, sxs, self.w13, self.s13, offsets, out_dtype=dt, rows_per_group_hint=hint)
        h = torch.nn.functional.silu(gu[:, :self.inter].float()).mul_(gu[:, self.inter:].float()).to(dt)
        L2, R2 = self.f2
        if self.diag2 is not None:
            h = h * self.diag2
        hq8, sh = ops.act_quant(h, L2, R2, L2.shape[0], R2.shape[0], *self.clip2)
        out = ops.gemm_w4a8_grouped(hq8, sh, self.w2, self.s2, offsets, out_dtype=dt, rows_per_group_hint=hint)
        w_sorted = weights.reshape(-1)[order].to(torch.float32)
        y = torch.zeros((T, self.dim), dtype=torch.float32, device=x.device)
        y.index_add_(0, tok, out.float() * w_sorted[:, None])
        return y.to(dt)
