"""Per-shape GPU time of the public W4A8 GEMM entry point (ops.gemm_w4a8), rotating over distinct weights so that
every launch streams its packed weight from HBM.  FLATQ_B200_LIB selects another build for A/B runs."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops

dev = torch.device("cuda:0")
shapes = [(2048, 6144, 4096), (2048, 4096, 4096), (2048, 28672, 4096), (2048, 4096, 14336), (8192, 4096, 4096), (512, 4096, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (m, n, k) in shapes:
    nw = max(2, min(8, int(400e6 // (n * k // 2))))
    x8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev)
    sx = torch.rand(m, device=dev) + 0.5
    ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(nw)]
    sw = torch.rand(n, device=dev) + 0.5
    for i in range(3):
        y = ops.gemm_w4a8(x8, sx, ws[i % nw], sw, None, torch.float16)
    torch.cuda.synchronize()
    reps = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        y = ops.gemm_w4a8(x8, sx, ws[i % nw], sw, None, torch.float16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"M={m} N={n} K={k}: {us:8.1f} us  {2 * m * n * k / us / 1e6:7.0f} TOPS")
