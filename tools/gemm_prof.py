"""Role-level cycle counters of the tcgen05 GEMM (debug entry point flatq_debug_gemm_profile)."""
import ctypes
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200 import _lib

lib = _lib.load()
fn = lib.flatq_debug_gemm_profile
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
dev = torch.device("cuda:0")
shapes = [(2048, 4096, 4096), (2048, 28672, 4096), (2048, 4096, 14336), (2048, 6144, 4096)]
NW = int(os.environ.get("FLATQ_PROF_WARPS", "28"))  # warps per CTA of the build under test
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split("x"))]
for (m, n, k) in shapes:
    x8 = torch.randint(-8, 8, (m, k), dtype=torch.int8, device=dev)
    sx = torch.rand(m, device=dev)
    w = torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev)
    sw = torch.rand(n, device=dev)
    y = torch.empty(m, n, dtype=torch.float16, device=dev)
    grid = min(148, ((n + 127) // 128) * ((m + 255) // 256))
    for dbg in [int(v) for v in os.environ.get("FLATQ_PROF_DBG", "0,4").split(",")]:
        prof = torch.zeros(grid, NW, 5, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            rc = fn(x8.data_ptr(), sx.data_ptr(), w.data_ptr(), sw.data_ptr(), m, n, k, y.data_ptr(), prof.data_ptr(), dbg, st)
            assert rc == 0, lib.flatq_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(x8.data_ptr(), sx.data_ptr(), w.data_ptr(), sw.data_ptr(), m, n, k, y.data_ptr(), prof.data_ptr(), dbg, st)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        p = prof.double().mean(0).cpu()  # [14, 5] averaged over CTAs
        kb_total = (k // 256) * ((n + 127) // 128) * ((m + 255) // 256) / grid  # k-blocks per CTA
        print(f"== M={m} N={n} K={k} dbg={dbg}: {us:.1f} us, {2*m*n*k/us/1e6:.0f} TOPS; kernel cycles {p[0,4]:.0f}; k-blocks/CTA {kb_total:.0f} -> {p[0,4]/kb_total:.0f} cyc/k-block")
        print(f"   producers: X wait empty {p[0,0]:.0f}   W wait wempty {p[2,0]:.0f}")
        print(f"   mma     : wait full {p[1,0]:.0f}  wait empty_acc {p[1,2]:.0f}")
        ex = p[NW - 24:NW - 16]
        print(f"   expander: wait wfull {ex[:,0].mean():.0f}  wait empty {ex[:,1].mean():.0f}  tmem st wait {ex[:,2].mean():.0f}  (per warp; each does half the stages)")
        ep = p[NW - 16:NW]
        print(f"   epilogue: wait full_acc {ep[:,0].mean():.0f}  busy {ep[:,1].mean():.0f}")
