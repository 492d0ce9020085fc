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
