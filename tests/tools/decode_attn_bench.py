"""Decode attention over the paged INT4 cache at Llama-3-8B head counts (32 query / 8 kv heads, head_dim 128), seq 2048:
GPU time per launch of (a) flatq_kv_decode_i4 (single pass, verified), (b) flatq_kv_decode_i4_split (split-KV, opt-in),
(c) the reference's FlashInferBatchDecodeKernel_i4 from oracle/_ref -- which has no grouped-query path, so it runs over
a cache holding the kv heads repeated to 32 (what the reference's cache stores, kv_cache.py:291-299).
CUDA events on the legacy default stream, 50 launches after 5 warm-ups; prints one JSON object per (batch, kernel).

    python tests/tools/decode_attn_bench.py            (first run of (b): see tests/test_zzz_decode_split_gpu.py)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flatquant_b200 import ops  # noqa: E402
from oracle import ref_kernels as R  # noqa: E402  (baseline leg only)

dev = torch.device("cuda:0")
HQ, HKV, D, SEQ, LAYERS = 32, 8, 128, 2048, 2


def cache(batch, heads, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    ps = SEQ
    kv_data = torch.zeros(batch, LAYERS, 2, heads, ps, D // 2, dtype=torch.uint8, device=dev)
    kv_param = torch.zeros(batch, LAYERS, 2, heads, ps, 2, dtype=torch.float16, device=dev)
    slot = (torch.arange(batch, device=dev)[:, None] * ps + torch.arange(SEQ, device=dev)[None]).reshape(-1).int()
    for sel in (0, 1):
        x = torch.randn(batch * SEQ, heads, D, device=dev, generator=g).half()
        ops.kv_quant(x, kv_data, kv_param, slot, 1, sel)
    ip = torch.arange(batch + 1, dtype=torch.int32, device=dev)
    ix = torch.arange(batch, dtype=torch.int32, device=dev)
    lp = torch.full((batch,), SEQ, dtype=torch.int32, device=dev)
    return kv_data, kv_param, ip, ix, lp


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for batch in (1, 16):
    c = cache(batch, HKV, 1)
    q = torch.randn(batch, HQ, D, device=dev).half()
    kv_bytes = batch * HKV * SEQ * 2 * (D // 2 + 4)
    rows = {"single-pass": lambda: ops.kv_decode_i4(q, *c, layer=1, split=False)}
    if os.environ.get("FLATQ_RUN_UNVERIFIED"):
        rows["split-kv"] = lambda: ops.kv_decode_i4(q, *c, layer=1, split=True, max_tokens_hint=SEQ)
    if R.available():
        c32 = cache(batch, HQ, 2)
        rows["reference FlashInfer (32 repeated heads)"] = lambda: R.batch_decode_i4(q, c32[0], c32[1], c32[2], c32[3], c32[4], 1)
    for name, fn in rows.items():
        try:
            us = timeit(fn)
            print(json.dumps({"batch": batch, "kernel": name, "us": us, "kv_bytes": kv_bytes,
                              "GBps_on_kv_heads": kv_bytes / us / 1e3}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"batch": batch, "kernel": name, "error": f"{type(e).__name__}: {e}"[:200]}), flush=True)
