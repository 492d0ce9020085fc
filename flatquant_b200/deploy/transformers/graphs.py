"""CUDA-graph replay of the model steps (streams and graphs instead of a tracing compiler).

A decode step of the real-quant Llama stack is ~10 small launches per layer behind ~40 Python-level torch calls: the
host needs ~0.5 ms per layer, the GPU a few tens of microseconds, so eager decode is host-bound by an order of
magnitude (the reference reports the same for its own stack: 0.47x decode "speed-up").  Every
kernel of the library launches on the caller's stream without host synchronisation or allocation,
so a step can be captured once and replayed:

  * `GraphedDecoder`  one-token steps over the paged INT4 cache.  The token position lives in a device tensor
    (`cache.pos_dev`): the KV slot, the attended length and the RoPE row are computed from it inside the graph, and it
    is incremented at the end of the captured step, so successive replays walk through the sequence.
  * `GraphedPrefill`  the fixed-shape prefill (batch x seq_len), whose launch arguments are all static.

Nothing here changes what is computed: the captured callables are the same Python functions eager mode runs
(`FlatQuantLlamaForCausalLM.decode_static` / `.forward`); tests compare graphed and eager outputs."""
from __future__ import annotations

import torch


class GraphedDecoder:
    def __init__(self, model, cache, warmup=2):
        self.model, self.cache = model, cache
        dev = cache.device
        cache.begin_static_decode()
        self.capacity = min(int(cache.page_size), int(cache.max_seq_len))   # static decode: one page per sequence
        self.tokens = torch.zeros((cache.batch_size, 1), dtype=torch.long, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # allocator / workspace warm-up; rewinds the position it advanced
                model.decode_static(self.tokens, cache)
                cache.pos_dev.sub_(1)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.logits = model.decode_static(self.tokens, cache)

    def step(self, tokens):
        """tokens [batch, 1] -> logits [batch, 1, vocab] (a static buffer, overwritten by the next step)."""
        # the captured step computes its KV slot and RoPE row on the device from cache.pos_dev and cannot grow the page
        # pool: past the capacity it would scatter into the next sequence's page (or past the allocation)
        if self.cache.length >= self.capacity:
            raise RuntimeError(f"GraphedDecoder.step: the static cache is full ({self.cache.length} of {self.capacity} "
                               "tokens per sequence); build the cache with a larger max_length")
        self.tokens.copy_(tokens, non_blocking=True)
        self.graph.replay()
        self.cache.length += 1
        return self.logits


class GraphedPrefill:
    def __init__(self, model, batch, seq_len, max_length=None, warmup=2, logits_to_keep=0):
        self.model = model
        dev = model.lm_head.weight.device
        self.cache = model.build_cache(batch, page_size=max(max_length or seq_len, seq_len), max_length=max_length or seq_len)
        self.ids = torch.zeros((batch, seq_len), dtype=torch.long, device=dev)
        self.seq_len, self.keep = seq_len, logits_to_keep
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.logits = self._run()

    def _run(self):
        self.cache.reset()
        return self.model(self.ids, past_key_values=self.cache, logits_to_keep=self.keep)[0]

    def __call__(self, ids):
        """ids [batch, seq_len] -> (logits, cache holding the prefilled sequences)."""
        self.ids.copy_(ids, non_blocking=True)
        self.graph.replay()
        self.cache.length = self.seq_len  # the host-side bookkeeping the captured forward did at capture time
        self.cache._needs_init = [False] * self.cache.n_layers
        self.cache._step_key = None
        return self.logits, self.cache
