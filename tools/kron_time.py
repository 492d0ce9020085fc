import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops
dev = torch.device("cuda:0")
for (n1, n2) in [(64, 64), (112, 128), (64, 128)]:
    m, k = 2048, n1 * n2
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    out = (torch.empty(m, k, dtype=torch.int8, device=dev), torch.empty(m, dtype=torch.float32, device=dev))
    for _ in range(3):
        ops.act_quant(x, L, R, n1, n2, 0.98, 0.98, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.act_quant(x, L, R, n1, n2, 0.98, 0.98, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"act_quant ({n1},{n2}) M={m}: {us:.1f} us  ({m * k * 3 / us / 1e3:.0f} GB/s of 2B in + 1B out)")
