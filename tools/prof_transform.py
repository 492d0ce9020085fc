"""A few launches of the transform + quantise kernel for ncu (tokens=2048, (64,64) and (112,128))."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops
dev = torch.device("cuda:0")
for (m, n1, n2) in [(2048, 64, 64), (2048, 112, 128)]:
    k = n1 * n2
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
    R = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
    out = (torch.empty(m, k, dtype=torch.int8, device=dev), torch.empty(m, dtype=torch.float32, device=dev))
    for i in range(3):
        ops.act_quant(x, L, R, n1, n2, 0.98, 0.98, out=out)
torch.cuda.synchronize()
