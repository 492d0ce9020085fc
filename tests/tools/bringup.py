"""GPU bring-up: run every kernel against the oracle and print diagnostics (not a test, not a bench).
Usage on the GPU box:  timeout 600 python tools/bringup.py [stage ...]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flatquant_b200 import ops  # noqa: E402
from oracle import flatq_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def stage_elementwise():
    q = rng.integers(-8, 8, (64, 256)).astype(np.int8)
    p = ops.pack_i4(torch.from_numpy(q).to(dev)).cpu().numpy()
    print("pack_i4 exact:", np.array_equal(p, O.pack_i4(q)))
    u = ops.unpack_i4(torch.from_numpy(O.pack_i4(q)).to(dev)).cpu().numpy()
    print("unpack_i4 exact:", np.array_equal(u, q.astype(np.int32)))
    x = (rng.standard_normal((32, 512)) * 3).astype(np.float16)
    s = (np.abs(x).max(1) / 7).astype(np.float16)
    qq = ops.sym_quant(torch.from_numpy(x).to(dev), torch.from_numpy(s).to(dev)).cpu().numpy()
    print("sym_quant exact:", np.array_equal(qq, O.pack_i4(O.sym_quant_half(x, s))))
    acc = rng.integers(-200000, 200000, (32, 64)).astype(np.int32)
    sr = rng.random(32).astype(np.float16)
    sc = (rng.random(64) * 0.01).astype(np.float16)
    y = ops.sym_dequant(torch.from_numpy(acc).to(dev), torch.from_numpy(sr).to(dev), torch.from_numpy(sc).to(dev)).cpu().numpy()
    ref = O.sym_dequant(acc, sr, sc).astype(np.float16)
    print("sym_dequant exact-mode equal:", np.array_equal(y, ref), "max diff", np.abs(y.astype(np.float32) - ref.astype(np.float32)).max())
    yk = ops.sym_dequant(torch.from_numpy(acc).to(dev), torch.from_numpy(sr).to(dev), torch.from_numpy(sc).to(dev), ref_kernel_arith=True).cpu().numpy()
    refk = O.sym_dequant_ref_kernel(acc, sr, sc)
    print("sym_dequant ref-kernel-mode equal:", np.array_equal(yk, refk), "n diff", int((yk != refk).sum()))


def gemm_case(m, n, k, verbose=True):
    a = rng.integers(-8, 8, (m, k)).astype(np.int8)
    b = rng.integers(-8, 8, (n, k)).astype(np.int8)
    c = ops.gemm_i4(torch.from_numpy(O.pack_i4(a)).to(dev), torch.from_numpy(O.pack_i4(b)).to(dev))
    torch.cuda.synchronize()
    c = c.cpu().numpy()
    ref = O.gemm_i4(a, b)
    ok = np.array_equal(c, ref)
    print(f"gemm_i4 M={m} N={n} K={k}: exact={ok}", flush=True)
    if not ok and verbose:
        bad = np.argwhere(c != ref)
        print("  mismatches:", len(bad), "of", c.size, "first:", bad[:5].tolist())
        print("  got ", c[:2, :8].tolist())
        print("  want", ref[:2, :8].tolist())
        print("  bad rows (m) uniq:", np.unique(bad[:, 0])[:16].tolist(), " bad cols (n) uniq:", np.unique(bad[:, 1])[:16].tolist())
    return ok


def gemm_probe():
    """one-hot probes: which (n,k)x(m,k') pairs contract -> reveals layout errors"""
    m, n, k = 16, 128, 128
    for kk in (0, 1, 2, 7, 8, 33, 127):
        a = np.zeros((m, k), np.int8); a[3, kk] = 1
        b = np.zeros((n, k), np.int8); b[5, :] = np.arange(k) % 7 + 1
        c = ops.gemm_i4(torch.from_numpy(O.pack_i4(a)).to(dev), torch.from_numpy(O.pack_i4(b)).to(dev)).cpu().numpy()
        nz = np.argwhere(c != 0)
        print(f"  probe k={kk}: expect C[3,5]={kk % 7 + 1}; nonzeros:", [(int(i), int(j), int(c[i, j])) for i, j in nz[:6]])


def stage_gemm():
    ok = gemm_case(16, 128, 128)
    if not ok:
        gemm_probe()
    for shp in [(128, 128, 128), (128, 128, 512), (128, 256, 4096), (100, 200, 192), (2048, 4096, 4096), (1, 4096, 4096), (33, 1024, 14336)]:
        gemm_case(*shp, verbose=not ok)


def stage_kron():
    for (n1, n2, tokens) in [(16, 16, 5), (64, 64, 7), (14, 16, 9), (112, 128, 3), (56, 64, 4)]:
        k = n1 * n2
        x = torch.randn(tokens, k).half()
        L = (torch.randn(n1, n1) / n1 ** 0.5).half()
        R = (torch.randn(n2, n2) / n2 ** 0.5).half()
        z = ops.kron_transform(x.to(dev), L.to(dev), R.to(dev), n1, n2).float().cpu()
        ref = O.kronecker_matmul(x.float(), L.float(), R.float())
        err = (z - ref).abs().max().item()
        print(f"kron {n1}x{n2}: max err {err:.4e} (max|z| {ref.abs().max():.3f})", flush=True)
        zl = ops.kron_transform(x.to(dev), L.to(dev), None, n1, n2).float().cpu()
        refl = torch.matmul(L.float().t(), x.float().reshape(-1, n1, n2)).reshape(tokens, k)
        zr = ops.kron_transform(x.to(dev), None, R.to(dev), n1, n2).float().cpu()
        refr = torch.matmul(x.float().reshape(-1, n1, n2), R.float()).reshape(tokens, k)
        print(f"   left-only err {(zl - refl).abs().max():.4e}  right-only err {(zr - refr).abs().max():.4e}")
        xq, s16, s32 = ops.transform_quant(x.to(dev), L.to(dev), R.to(dev), n1, n2, 0.98, 0.95)
        codes = O.unpack_i4(xq.cpu().numpy())
        # oracle on the kernel's own fp16 z (isolates the quantiser) and end-to-end
        zk = ops.kron_transform(x.to(dev), L.to(dev), R.to(dev), n1, n2).cpu()
        hi = zk.float().amax(1).clamp(min=0) * 0.98
        lo = zk.float().amin(1).clamp(max=0) * 0.95
        s_ref = torch.maximum(lo.abs(), hi) / 7
        s_ref[s_ref == 0] = 1
        q_ref = torch.clamp(torch.round(zk.float() / s_ref[:, None]), -8, 7).numpy().astype(np.int32)
        print(f"   quant: scale equal {torch.equal(s32.cpu(), s_ref)}; codes equal (given kernel z) {np.array_equal(codes, q_ref)}"
              f" mismatches {int((codes != q_ref).sum())}")


def stage_fused():
    for (m, n, n1, n2) in [(96, 256, 64, 64), (2048, 4096, 64, 64), (7, 128, 16, 16)]:
        k = n1 * n2
        x = torch.randn(m, k).half()
        L = (torch.randn(n1, n1) / n1 ** 0.5).half()
        R = (torch.randn(n2, n2) / n2 ** 0.5).half()
        wq = torch.randint(-8, 8, (n, k), dtype=torch.int8)
        s_w = (torch.rand(n) * 0.02 + 0.01).float()
        cm = float(torch.sigmoid(torch.tensor(4.0)))
        wp = torch.from_numpy(O.pack_i4(wq.numpy()))
        y = ops.linear_fused(x.to(dev), L.to(dev), R.to(dev), n1, n2, cm, cm, wp.to(dev), s_w.to(dev))
        torch.cuda.synchronize()
        ref = O.flat_linear_int(x.float(), L.float(), R.float(), 4.0, 4.0, wq.numpy(), s_w.numpy())["y"]
        err = np.abs(y.float().cpu().numpy() - ref)
        print(f"fused M={m} N={n} K={k}: max|y| {np.abs(ref).max():.3f} max err {err.max():.4f} mean err {err.mean():.5f}", flush=True)


def stage_time():
    m, n, k = 2048, 4096, 4096
    a = torch.randint(0, 255, (m, k // 2), dtype=torch.uint8, device=dev)
    b = torch.randint(0, 255, (n, k // 2), dtype=torch.uint8, device=dev)
    sa = torch.rand(m, device=dev).half()
    sb = torch.rand(n, device=dev).half()
    for name, fn in [("gemm_i4 (expand+gemm, int32 out)", lambda: ops.gemm_i4(a, b)),
                     ("linear4bit (expand+gemm+dequant)", lambda: ops.linear4bit(a, sa, b, sb))]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name}: {ms * 1e3:.1f} us  -> {2 * m * n * k / ms / 1e9:.1f} TOPS", flush=True)
    x = torch.randn(m, k, device=dev).half()
    L = (torch.randn(64, 64, device=dev) / 8).half()
    R = (torch.randn(64, 64, device=dev) / 8).half()
    wp = b
    s_w = torch.rand(n, device=dev)
    for name, fn in [("transform_quant 64x64", lambda: ops.transform_quant(x, L, R, 64, 64)),
                     ("linear_fused", lambda: ops.linear_fused(x, L, R, 64, 64, 1.0, 1.0, wp, s_w))]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)


STAGES = {"elementwise": stage_elementwise, "gemm": stage_gemm, "kron": stage_kron, "fused": stage_fused, "time": stage_time}

if __name__ == "__main__":
    names = sys.argv[1:] or list(STAGES)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for nm in names:
        print(f"===== {nm}", flush=True)
        t0 = time.time()
        try:
            STAGES[nm]()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("device error after stage:", e)
                break
        print(f"===== {nm} done in {time.time() - t0:.1f}s", flush=True)
