"""Decode-shaped FlatLinears (1 / 16 / 64 tokens): GPU time per fused linear (transform+quantise+GEMM) measured inside a
CUDA graph (host launch overhead excluded), against the HBM floor of streaming the packed weight once."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops
dev = torch.device("cuda:0")
shapes = [(4096, 6144, 64, 64), (4096, 4096, 64, 64), (4096, 28672, 64, 64), (14336, 4096, 112, 128)]
for m in (1, 16, 64):
    for (k, n, n1, n2) in shapes:
        x = torch.randn(m, k, device=dev).half()
        L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
        R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
        ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(16)]  # 16 distinct weights
        sw = torch.rand(n, device=dev)
        xq = (torch.empty(m, k, dtype=torch.int8, device=dev), torch.empty(m, dtype=torch.float32, device=dev))
        y = torch.empty(m, n, dtype=torch.float16, device=dev)
        def run(i):
            ops.act_quant(x, L, R, n1, n2, 0.98, 0.98, out=xq)
            ops.gemm_w4a8(xq[0], xq[1], ws[i % 16], sw, out=y)
        for i in range(4):
            run(i)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(2):
                run(i)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(16):
                    run(i)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 80 * 1e3
        floor = n * k / 2 / 6.57e6  # us at 6.57 TB/s
        print(f"M={m:3d} {k}->{n}: {us:7.1f} us per linear in a CUDA graph (weight stream floor {floor:5.1f} us -> {floor / us * 100:4.1f}% of HBM roofline)")
