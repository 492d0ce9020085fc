import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from flatquant_b200 import ops
from oracle import flatq_oracle as O
dev = torch.device('cuda:0')
for trial in range(4):
    m, n, k = 2048, 4096, 4096
    rng = np.random.default_rng(m * 7 + n + trial)
    a = rng.integers(-8, 8, (m, k)).astype(np.int8)
    b = rng.integers(-8, 8, (n, k)).astype(np.int8)
    if trial % 2 == 0:
        a[0, :] = -8; b[0, :] = -8
    c = ops.gemm_i4(torch.from_numpy(O.pack_i4(a)).to(dev), torch.from_numpy(O.pack_i4(b)).to(dev)).cpu().numpy()
    ref = O.gemm_i4(a, b)
    bad = np.argwhere(c != ref)
    print("trial", trial, "extreme" if trial % 2 == 0 else "random", "mismatches", len(bad))
    if len(bad):
        print(" first", bad[:6].tolist(), "got", c[tuple(bad[0])], "want", ref[tuple(bad[0])])
        print(" bad rows uniq", np.unique(bad[:, 0])[:12].tolist(), "bad cols uniq", np.unique(bad[:, 1])[:12].tolist(), "n bad cols", len(np.unique(bad[:,1])))
