"""The decode-shape kernels once each per Llama-3-8B linear, for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:'skinny|flatlinear_decode|gemv_w4a8' -o out python tools/prof_decode_kernels.py
16 tokens: gemm_w4a8_skinny_kernel (qkv, o, gate/up + SiLU, down); 1 token: flatlinear_decode_kernel (qkv, o, gate/up + SiLU) and
gemv_w4a8_pre_kernel (down).  Distinct weights per launch: every launch streams from HBM."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops

dev = torch.device("cuda:0")
shapes = (("qkv", 4096, 6144, 64, 64, False), ("o", 4096, 4096, 32, 128, True), ("gate_up", 4096, 28672, 64, 64, False),
          ("down", 14336, 4096, 112, 128, False))
for m in (16, 1):
    for name, k, n, n1, n2, left_only in shapes:
        x = torch.randn(m, k, device=dev).half()
        L = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
        R = None if left_only else (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
        ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(3)]
        sw = torch.rand(n, device=dev) * 0.01 + 0.005
        for w in ws:
            if name == "gate_up":
                ops.linear_fused_silu(x, L, R, n1, n2, 0.98, 0.98, w, sw)
            else:
                ops.linear_fused(x, L, R, n1, n2, 0.98, 0.98, w, sw)
torch.cuda.synchronize()
