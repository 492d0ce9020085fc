import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import ops
dev = torch.device("cuda:0")
D, HQ, HKV, L = 128, 32, 8, 1
for tokens in (2048, 1):
    b, s = 1, tokens
    qkv = torch.randn(tokens, (HQ + 2 * HKV) * D, device=dev).half()
    q, k, v = (t.reshape(b, s, -1, D) for t in torch.split(qkv.view(b, s, -1), [HQ * D, HKV * D, HKV * D], dim=-1))
    ang = torch.arange(4096, device=dev)[:, None].float() * (1.0 / 500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))[None]
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = emb.cos().half(), emb.sin().half()
    pos = torch.arange(s, device=dev).int()
    t_k = (torch.eye(D, device=dev) * 0.75).half()
    ps = 2048 + 64
    pages = torch.zeros(b, L, 2, HKV, ps, D // 2, dtype=torch.uint8, device=dev)
    params = torch.zeros(b, L, 2, HKV, ps, 2, dtype=torch.float16, device=dev)
    slot = torch.arange(s, device=dev).int()
    for _ in range(3):
        ops.kv_fused(q, k, v, cos_t, sin_t, pos, t_k, t_k, pages, params, slot, 0, want_k=True)
torch.cuda.synchronize()
