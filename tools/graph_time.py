"""GPU-side time of (a) the transform + quantise launch alone and (b) one fused FlatLinear, per shape, replayed from a CUDA
graph of 16 launches over distinct inputs / weights -- the host's per-call cost (ctypes, tensor-map encoding, allocator:
~10-20 us, more than the small kernels themselves) is out of the measurement.  FLATQ_KRON_TC1=1: first tcgen05 transform."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops

dev = torch.device("cuda:0")
peak = 2 * json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]


def graph_us(fn, n=16, reps=5):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                fn(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n) * 1e3


which = sys.argv[1:] or ["transform", "fused"]
if "transform" in which:
    for (m, n1, n2) in [(2048, 64, 64), (2048, 112, 128), (2048, 64, 128), (2048, 56, 64), (2048, 32, 64), (2048, 128, 128), (8192, 64, 64), (16, 64, 64), (1, 64, 64)]:
        k = n1 * n2
        xs = [torch.randn(m, k, device=dev).half() for _ in range(16)]
        L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
        R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
        out = (torch.empty(m, k, dtype=torch.int8, device=dev), torch.empty(m, dtype=torch.float32, device=dev))
        us = graph_us(lambda i: ops.act_quant(xs[i % 16], L, R, n1, n2, 0.98, 0.98, out=out))
        print(f"transform tokens={m} ({n1},{n2}) K={k}: {us:7.2f} us  {3.0 * m * k / us / 1e3:7.0f} GB/s (2 B in + 1 B out per element)", flush=True)
if "fused" in which:
    m = 2048
    for (k, n, n1, n2) in [(4096, 4096, 64, 64), (4096, 6144, 64, 64), (4096, 28672, 64, 64), (14336, 4096, 112, 128)]:
        x = torch.randn(m, k, device=dev).half()
        L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
        R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
        ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(16)]
        sw = torch.rand(n, device=dev) * 0.01
        us = graph_us(lambda i: ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, ws[i % 16], sw))
        xq = ops.act_quant(x, L, R, n1, n2, 0.98, 0.98)
        y = torch.empty(m, n, dtype=torch.float16, device=dev)
        us_g = graph_us(lambda i: ops.gemm_w4a8(xq[0], xq[1], ws[i % 16], sw, out=y))
        tops = 2 * m * n * k / us / 1e6
        print(f"fused linear 2048 x {k}->{n}: {us:7.1f} us (GEMM alone {us_g:6.1f})  {tops:7.0f} TOP/s = {100 * tops / peak:4.1f}% of measured burst peak ({peak:.0f})", flush=True)
if "side" in which:
    # the launches between the FlatLinears of one decoder layer, replayed from a graph (decode bsz 1 / 16 and prefill 2048)
    D, HQ, HKV, L = 128, 32, 8, 1
    for tokens, seq in ((1, 2048), (16, 2048), (2048, 0)):
        b = tokens if seq else 1
        s = 1 if seq else tokens
        qkv = torch.randn(tokens, (HQ + 2 * HKV) * D, device=dev).half()
        q, k, v = (t.reshape(b, s, -1, D) for t in torch.split(qkv.view(b, s, -1), [HQ * D, HKV * D, HKV * D], dim=-1))
        P = 4096
        ang = torch.arange(P, device=dev)[:, None].float() * (1.0 / 500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))[None]
        emb = torch.cat((ang, ang), -1)
        cos_t, sin_t = emb.cos().half(), emb.sin().half()
        pos = (torch.arange(s, device=dev) + (seq or 0)).repeat(b).int()
        t_k = (torch.eye(D, device=dev) * 0.75 + torch.randn(D, D, device=dev) * (0.5 / D ** 0.5)).half()
        t_q = torch.linalg.inv(t_k.double()).t().float().half().contiguous()
        ps = 2048 + 64
        pages = torch.zeros(b, L, 2, HKV, ps, D // 2, dtype=torch.uint8, device=dev)
        params = torch.zeros(b, L, 2, HKV, ps, 2, dtype=torch.float16, device=dev)
        slot = (torch.arange(b, device=dev)[:, None] * ps + (torch.arange(s, device=dev) + (seq or 0))[None]).reshape(-1).int()
        us = graph_us(lambda i: ops.kv_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, pages, params, slot, 0, want_k=not seq))
        print(f"kv_fused tokens={tokens}: {us:7.2f} us", flush=True)
        h = torch.randn(tokens, 4096, device=dev).half()
        w = torch.rand(4096, device=dev).half()
        us = graph_us(lambda i: ops.add_rmsnorm(h, h, w, 1e-5))
        print(f"add_rmsnorm tokens={tokens}: {us:7.2f} us", flush=True)
        gu = torch.randn(tokens, 2 * 14336, device=dev).half()
        us = graph_us(lambda i: ops.silu_mul(gu))
        print(f"silu_mul tokens={tokens}: {us:7.2f} us", flush=True)
        if seq:
            kx = torch.randn(b * seq, HKV, D, device=dev).half()
            sl = (torch.arange(b, device=dev)[:, None] * ps + torch.arange(seq, device=dev)[None]).reshape(-1).int()
            ops.kv_quant(kx, pages, params, sl, 0, 0)
            ops.kv_quant(kx, pages, params, sl, 0, 1)
            ip = torch.arange(b + 1, dtype=torch.int32, device=dev)
            ix = torch.arange(b, dtype=torch.int32, device=dev)
            lp = torch.full((b,), seq + 1, dtype=torch.int32, device=dev)
            qq = torch.randn(b, HQ, D, device=dev).half()
            us = graph_us(lambda i: ops.kv_decode_i4(qq, pages, params, ip, ix, lp, layer=0, max_tokens_hint=ps))
            kvb = 2.0 * b * HKV * (seq + 1) * 68
            print(f"decode attention (split + merge) bsz={b} seq={seq}: {us:7.2f} us  {kvb / us / 1e3:6.0f} GB/s", flush=True)
            for (kk, nn, n1, n2, left_only) in ((4096, 6144, 64, 64, False), (4096, 4096, 32, 128, True), (4096, 28672, 64, 64, False), (14336, 4096, 112, 128, False)):
                x = torch.randn(tokens, kk, device=dev).half()
                Lm = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
                Rm = None if left_only else (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
                ws = [torch.randint(0, 256, (nn, kk // 2), dtype=torch.uint8, device=dev) for _ in range(16)]
                sw = torch.rand(nn, device=dev) * 0.01
                out = (torch.empty(tokens, kk, dtype=torch.int8, device=dev), torch.empty(tokens, dtype=torch.float32, device=dev))
                y = torch.empty(tokens, nn, dtype=torch.float16, device=dev)
                us_t = graph_us(lambda i: ops.act_quant(x, Lm, Rm, n1, n2, 0.98, 0.98, out=out))
                us_g = graph_us(lambda i: ops.gemm_w4a8(out[0], out[1], ws[i % 16], sw, out=y))
                us_c = graph_us(lambda i: ops.gemm_w4a8(*ops.act_quant(x, Lm, Rm, n1, n2, 0.98, 0.98, out=out), ws[i % 16], sw, out=y))
                extra = ""
                if nn == 28672:
                    ya = torch.empty(tokens, nn // 2, dtype=torch.float16, device=dev)
                    us_s = graph_us(lambda i: ops.gemm_w4a8_silu(out[0], out[1], ws[i % 16], sw, out=ya))
                    us_m = graph_us(lambda i: ops.silu_mul(y))
                    extra = f", GEMM+SiLU in one launch {us_s:6.2f} us (separate silu_mul {us_m:5.2f} us)"
                print(f"  M={tokens} {kk}->{nn}: transform {us_t:6.2f} us, GEMV/GEMM {us_g:6.2f} us, chained (PDL) {us_c:6.2f} us "
                      f"(weight stream floor {kk * nn / 2 / 6.544e6:5.2f} us){extra}", flush=True)

if "kvf" in which:
    # kv_fused alone over decode batch sizes, with parts switched off (what scales with the number of tokens?)
    D, HQ, HKV, L = 128, 32, 8, 1
    P = 4096
    ang = torch.arange(P, device=dev)[:, None].float() * (1.0 / 500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))[None]
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = emb.cos().half(), emb.sin().half()
    t_k = (torch.eye(D, device=dev) * 0.75 + torch.randn(D, D, device=dev) * (0.5 / D ** 0.5)).half()
    t_q = torch.linalg.inv(t_k.double()).t().float().half().contiguous()
    ps = 2048 + 64
    for b in (1, 2, 4, 8, 16, 32):
        qkv = torch.randn(b, (HQ + 2 * HKV) * D, device=dev).half()
        q, k, v = (t.reshape(b, 1, -1, D) for t in torch.split(qkv.view(b, 1, -1), [HQ * D, HKV * D, HKV * D], dim=-1))
        pos = torch.full((b,), 2048, device=dev).int()
        pages = torch.zeros(b, L, 2, HKV, ps, D // 2, dtype=torch.uint8, device=dev)
        params = torch.zeros(b, L, 2, HKV, ps, 2, dtype=torch.float16, device=dev)
        slot = (torch.arange(b, device=dev) * ps + 2048).int()
        full = graph_us(lambda i: ops.kv_fused(q, k, v, cos_t, sin_t, pos, t_q, t_k, pages, params, slot, 0))
        norope = graph_us(lambda i: ops.kv_fused(q, k, v, None, None, None, t_q, t_k, pages, params, slot, 0))
        not_ = graph_us(lambda i: ops.kv_fused(q, k, v, cos_t, sin_t, pos, None, None, pages, params, slot, 0))
        print(f"kv_fused bsz={b}: {full:6.2f} us, without RoPE {norope:6.2f}, without transforms {not_:6.2f}", flush=True)
