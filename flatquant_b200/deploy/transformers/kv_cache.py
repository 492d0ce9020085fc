"""Paged 4-bit KV cache (mirror of deploy/transformers/kv_cache.py:170-359, without the HF `Cache` base class so that it
does not depend on a transformers version).  Same constructor arguments, same storage layout

    pages  [max_pages * batch, n_layers, 2, heads, page_size, head_dim/2] u8
    scales [max_pages * batch, n_layers, 2, heads, page_size, 2] f16   (scale, zero)

and the same `update(key_states, value_states, layer_idx, cache_kwargs)` protocol: the first call per layer (prefill)
stores the quantised K/V and returns the (transformed) fp16 states for flash attention; later calls (one token) append
and return a closure that runs single-query attention over the packed cache.

Per call the reference launches ~15 torch kernels + FHT/matmul + an append kernel (kv_cache.py:263-353); here it is
one transform launch (Hadamard or 128x128 matmul) + one fused quantise/pack/scatter launch per K and V, and one
decode-attention launch."""
import functools
import math

import torch

from ... import ops


class MultiLayerPagedKVCache4Bit:
    def __init__(self, batch_size, page_size, max_seq_len, device, n_layers, num_heads, head_dim, disable_quant=False,
                 trans_dtype=torch.float16, trans="had", group_size=1):
        if disable_quant:
            raise NotImplementedError("flatquant_b200 implements the INT4 cache only")
        self.batch_size, self.page_size, self.max_seq_len = batch_size, page_size, max_seq_len
        self.device, self.n_layers, self.num_heads, self.org_head_dim = device, n_layers, num_heads, head_dim
        self.trans, self.trans_dtype, self.group_size = trans, trans_dtype, group_size
        self.max_page_cnt = self.page_cnt_from_length(max_seq_len)
        self.pages = torch.zeros((self.max_page_cnt * batch_size, n_layers, 2, num_heads, page_size, head_dim // 2),
                                 dtype=torch.uint8, device=device)
        self.scales = torch.zeros((self.max_page_cnt * batch_size, n_layers, 2, num_heads, page_size, 2),
                                  dtype=torch.float16, device=device)
        self._needs_init = [True] * n_layers
        self.length = 0
        # sequence b owns pages [b * max_page_cnt, (b + 1) * max_page_cnt)
        self._page_base = torch.arange(batch_size, dtype=torch.int32, device=device) * self.max_page_cnt

    def reset(self):
        """Forget the cached sequences (the page pool is kept): the next update per layer is a prefill again."""
        self._needs_init = [True] * self.n_layers
        self.length = 0
        self._step_key = None

    def page_cnt_from_length(self, length):
        return (length + self.page_size - 1) // self.page_size

    @property
    def seen_tokens(self):
        return self.length

    def get_seq_length(self, layer_idx=0) -> int:
        """Length of the cached sequences (kv_cache.py:378-380)."""
        return self.length

    def get_usable_length(self, new_seq_length=0, layer_idx=0) -> int:
        return self.length

    def get_max_length(self):
        """No hard maximum: the page pool grows on demand (kv_cache.py:382-384)."""
        return None

    def to_legacy_cache(self):
        return self

    def _ensure_page_cnt_per_batch(self, expected):
        """Grow the page pool (at least doubling, kv_cache.py:212-231).  Sequence b keeps the contiguous page range
        [b * max_page_cnt, (b + 1) * max_page_cnt), so the pool is re-laid out with the new per-sequence count."""
        if expected <= self.max_page_cnt:
            return
        new_cnt = max(expected, 2 * self.max_page_cnt)
        for name in ("pages", "scales"):
            old = getattr(self, name)
            grown = torch.zeros((self.batch_size, new_cnt) + tuple(old.shape[1:]), dtype=old.dtype, device=old.device)
            grown[:, :self.max_page_cnt] = old.view((self.batch_size, self.max_page_cnt) + tuple(old.shape[1:]))
            setattr(self, name, grown.view((self.batch_size * new_cnt,) + tuple(old.shape[1:])))
        self.max_page_cnt = new_cnt
        self.max_seq_len = new_cnt * self.page_size
        self._page_base = torch.arange(self.batch_size, dtype=torch.int32, device=self.device) * new_cnt

    def get_cache_specs_for_flash_infer(self, attention_mask=None):
        """Page-table view of the cache under the reference's argument names (kv_cache.py:356-376)."""
        if attention_mask is not None:
            raise NotImplementedError("padded batches: all sequences of a batch share one length here")
        kw = self._attention_kwargs()
        return dict(kv_data=self.pages, kv_param=self.scales, kv_indptr=kw["kv_indptr"], kv_indices=kw["kv_indices"],
                    last_page_offset=kw["last_page_len"])

    def _transform(self, x, matrix):
        """K (and Q at decode) <- x . T_k: 128x128 matmul ('matmul*') or H_128/sqrt(128) ('had')."""
        if self.trans == "had":
            return ops.hadamard(x.to(self.trans_dtype).contiguous(), 1.0 / math.sqrt(x.shape[-1]))
        if self.trans.startswith("matmul"):
            flat = x.to(self.trans_dtype).reshape(-1, x.shape[-1]).contiguous()
            n = flat.shape[-1]
            m = matrix.to(flat).contiguous()
            # right-only configuration (x . T, no left factor): hardware-verified in round 2
            # (tests/test_zz_llama_model_gpu.py::test_right_only_transform_and_matmul_kv_cache)
            return ops.kron_transform(flat, None, m, 1, n).view(x.shape)
        return x

    def _slots(self, start, count):
        """slot = page * page_size + offset for tokens [start, start + count) of every sequence -> int32 [batch * count]"""
        t = torch.arange(start, start + count, dtype=torch.int32, device=self.device)
        page = self._page_base[:, None] + (t // self.page_size)[None, :]
        return (page * self.page_size + (t % self.page_size)[None, :]).reshape(-1).contiguous()

    def _attention_kwargs(self):
        n_pages = self.page_cnt_from_length(self.length)
        indptr = torch.arange(self.batch_size + 1, dtype=torch.int32, device=self.device) * n_pages
        indices = (self._page_base[:, None] + torch.arange(n_pages, dtype=torch.int32, device=self.device)[None, :]).reshape(-1)
        last = torch.full((self.batch_size,), self.length - (n_pages - 1) * self.page_size, dtype=torch.int32, device=self.device)
        return dict(kv_indptr=indptr, kv_indices=indices.contiguous(), last_page_len=last)

    def _decode(self, q, num_kv_heads, attention_kwargs, layer_idx, k_inv_t):
        b, q_len, hq, d = q.shape
        assert q_len == 1
        qt = self._transform(q.view(b, hq, d), k_inv_t)
        # the split-KV kernel (opt-in) sizes its grid from a host-side bound of the length; in static (graph) mode the
        # length lives on the device, so the bound is the cache capacity
        bound = self.max_seq_len if getattr(self, "_static_mode", False) else self.length
        o = ops.kv_decode_i4(qt.contiguous(), self.pages, self.scales, layer=layer_idx, max_tokens_hint=bound,
                             **attention_kwargs)
        return o.unsqueeze(1)

    # ---- decode with the position held on the device: every launch argument of a step is then either static or a
    # device tensor updated in place, so the whole step can be captured in a CUDA graph and replayed (graphs.py)
    def begin_static_decode(self):
        if self.max_page_cnt != 1:
            raise RuntimeError("static (graph-capturable) decode needs one page per sequence: page_size >= max_seq_len, "
                               "which is how the reference sizes its cache too (modeling_llama.py:337)")
        b, dev = self.batch_size, self.device
        self.pos_dev = torch.full((1,), self.length, dtype=torch.int32, device=dev)  # tokens cached so far
        self._static_attn = dict(kv_indptr=torch.arange(b + 1, dtype=torch.int32, device=dev),
                                 kv_indices=self._page_base.clone(),
                                 last_page_len=torch.empty(b, dtype=torch.int32, device=dev))
        self._static_slots = torch.empty(b, dtype=torch.int32, device=dev)
        self._static_mode = True

    def end_static_decode(self):
        self._static_mode = False

    def _update_static(self, key_states, value_states, layer_idx, cache_kwargs):
        k_mat, k_inv_t = cache_kwargs.get("trans_matrix_k"), cache_kwargs.get("trans_matrix_k_inv_t")
        b, added, heads, d = key_states.shape
        assert b == self.batch_size and added == 1
        key_t = self._transform(key_states, k_mat)
        if layer_idx == 0:  # slot of the new token and the attended length, from the device-side position
            self._static_slots.copy_(self._page_base * self.page_size + self.pos_dev)
            self._static_attn["last_page_len"].copy_((self.pos_dev + 1).expand(b))
        kq = key_t.reshape(b, heads, d).to(torch.float16).contiguous()
        vq = value_states.reshape(b, heads, d).to(torch.float16).contiguous()
        ops.kv_quant(kq, self.pages, self.scales, self._static_slots, layer_idx, 0)
        ops.kv_quant(vq, self.pages, self.scales, self._static_slots, layer_idx, 1)
        return functools.partial(self._decode, num_kv_heads=heads, attention_kwargs=self._static_attn,
                                 layer_idx=layer_idx, k_inv_t=k_inv_t if self.trans.startswith("matmul") else None)

    def fused_ok(self, q):
        """Can this step go through ops.kv_fused (one launch for RoPE, K / Q transform, KV quantise, page scatter)?"""
        return (ops.FUSED_KV and self.org_head_dim == 128 and q.dtype == torch.float16 and q.shape[-1] == 128
                and (self.trans == "none" or self.trans.startswith("matmul")))

    def fused_update(self, query_states, key_states, value_states, layer_idx, rope, cache_kwargs=None):
        """`update` for the fused path: takes the PRE-RoPE q / k / v (views of the qkv GEMM output are read in place) and
        the RoPE context `rope` = (cos table [P, 128], sin table, pos int32 [tokens]); returns
        prefill: (q_t, k_t, v) -- the transformed fp16 states the causal attention runs on;
        decode:  the attention output [b, 1, Hq, 128] over the packed INT4 cache (the closure of `update`, applied)."""
        cache_kwargs = cache_kwargs or {}
        matmul = self.trans.startswith("matmul")
        t_k = cache_kwargs.get("trans_matrix_k") if matmul else None
        t_q = cache_kwargs.get("trans_matrix_k_inv_t") if matmul else None
        b, added, heads, d = key_states.shape
        assert b == self.batch_size
        cos_t, sin_t, pos = rope
        static = getattr(self, "_static_mode", False)
        if static:
            assert added == 1
            if layer_idx == 0:
                self._static_slots.copy_(self._page_base * self.page_size + self.pos_dev)
                self._static_attn["last_page_len"].copy_((self.pos_dev + 1).expand(b))
            slots, prefill = self._static_slots, False
        else:
            if layer_idx == 0:
                self._ensure_page_cnt_per_batch(self.page_cnt_from_length(self.length + added))
                self._start = self.length
                self.length += added
            key = (self._start, added, self.max_page_cnt)
            if getattr(self, "_step_key", None) != key:
                self._step_key, self._step_slots, self._step_attn = key, self._slots(self._start, added), None
            slots = self._step_slots
            prefill = self._needs_init[layer_idx]
            self._needs_init[layer_idx] = False
        hq = query_states.shape[-2]
        if not prefill:
            assert added == 1
            if static:
                kw, bound = self._static_attn, self.max_seq_len
            else:
                if self._step_attn is None:
                    self._step_attn = self._attention_kwargs()
                kw, bound = self._step_attn, self.length
            if ops.kv_decode_fused_ok(query_states, key_states, cos_t, t_q, t_k):
                # the whole step between the qkv GEMM and o_proj in one kernel (+ the split merge): the attention kernel's
                # prologue does what kv_fused does for the one new token per sequence
                o = ops.kv_decode_fused(query_states, key_states, value_states, cos_t, sin_t, pos, t_q, t_k, self.pages,
                                        self.scales, slots, layer=layer_idx, max_tokens_hint=bound, **kw)
                return o.unsqueeze(1)
        q_t, k_t = ops.kv_fused(query_states, key_states, value_states, cos_t, sin_t, pos, t_q, t_k, self.pages, self.scales,
                                slots, layer_idx, want_k=prefill)
        if prefill:
            return q_t.view(b, added, hq, d), k_t.view(b, added, heads, d), value_states
        o = ops.kv_decode_i4(q_t, self.pages, self.scales, layer=layer_idx, max_tokens_hint=bound, **kw)
        return o.unsqueeze(1)

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        cache_kwargs = cache_kwargs or {}
        if getattr(self, "_static_mode", False):
            return self._update_static(key_states, value_states, layer_idx, cache_kwargs)
        k_mat, k_inv_t = cache_kwargs.get("trans_matrix_k"), cache_kwargs.get("trans_matrix_k_inv_t")
        b, added, heads, d = key_states.shape
        assert b == self.batch_size
        key_t = self._transform(key_states, k_mat)
        if layer_idx == 0:
            self._ensure_page_cnt_per_batch(self.page_cnt_from_length(self.length + added))
            self._start = self.length
            self.length += added
        # slot indices and the page table are the same for every layer of a step: build them once, not per layer
        key = (self._start, added, self.max_page_cnt)
        if getattr(self, "_step_key", None) != key:
            self._step_key, self._step_slots, self._step_attn = key, self._slots(self._start, added), None
        slots = self._step_slots
        kq = key_t.reshape(b * added, heads, d).to(torch.float16).contiguous()
        vq = value_states.reshape(b * added, heads, d).to(torch.float16).contiguous()
        ops.kv_quant(kq, self.pages, self.scales, slots, layer_idx, 0)
        ops.kv_quant(vq, self.pages, self.scales, slots, layer_idx, 1)
        if self._needs_init[layer_idx]:
            self._needs_init[layer_idx] = False
            return key_t, value_states  # prefill: attention runs on the (transformed) fp16 states
        assert added == 1
        if self._step_attn is None:
            self._step_attn = self._attention_kwargs()
        return functools.partial(self._decode, num_kv_heads=heads, attention_kwargs=self._step_attn,
                                 layer_idx=layer_idx, k_inv_t=k_inv_t if self.trans.startswith("matmul") else None)


class HFCacheAdapter:
    """Forwarding wrapper with the `Cache` protocol the reference hands to HF's LlamaModel (kv_cache.py:390-412)."""

    def __init__(self, inner):
        self.inner = inner

    def __getattr__(self, name):  # page pool, trans mode, ... are read straight from the wrapped cache
        return getattr(self.inner, name)

    def get_usable_length(self, kv_seq_len=0, layer_idx=0) -> int:
        return self.inner.get_usable_length(kv_seq_len, layer_idx)

    def get_seq_length(self, layer_idx=0) -> int:
        return self.inner.get_seq_length()

    def get_max_length(self):
        return self.inner.get_max_length()

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        return self.inner.update(key_states, value_states, layer_idx, cache_kwargs)

    def to_legacy_cache(self):
        return self.inner.to_legacy_cache()
