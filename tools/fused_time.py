"""GPU time of one fused FlatQuantizedLinear (transform + quantise + W4A4 GEMM + dequant) at bsz*seq = 2048, through
ops.linear_fused, rotating over distinct weights; reports the fraction of the measured INT4(int8-path) tensor peak."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops

dev = torch.device("cuda:0")
peak = 2 * 1721.2
try:
    peak = 2 * json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
m = 2048
for (k, n, n1, n2) in [(4096, 4096, 64, 64), (4096, 6144, 64, 64), (4096, 28672, 64, 64), (14336, 4096, 112, 128)]:
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    nw = max(2, min(8, int(400e6 // (n * k // 2))))
    ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(nw)]
    sw = torch.rand(n, device=dev) * 0.01
    bias = torch.randn(n, device=dev).half()
    for i in range(4):
        ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, ws[i % nw], sw, bias)
    torch.cuda.synchronize()
    reps = 32
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, ws[i % nw], sw, bias)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    tops = 2 * m * n * k / us / 1e6
    print(f"fused linear 2048 x {k}->{n}: {us:7.1f} us  {tops:7.0f} TOP/s = {100 * tops / peak:4.1f}% of measured peak ({peak:.0f}), {100 * tops / 4500:4.1f}% of nominal 4.5 POP/s")
