"""BASELINE LEG (not the product): the reference's OWN native GEMM + dequant kernels -- CUTLASS Sm80 s4 GEMM
(deploy/kernels/gemm.cu:7-46) and sym_dequant (quant.cu:66-101), compiled for sm_100a from /root/reference into
oracle/_ref (see oracle/ref_kernels_shim.cu) -- timed on this GPU over the same work as bench.py's GEMM-only roofline
pass: 32 layers x {qkv 4096->6144, o 4096->4096, gate_up 4096->28672, down 14336->4096} at 2048 tokens, distinct packed
weights per layer (3.5 GB, larger than L2).  This is the "reference's own deploy/kernels CUTLASS build on the same box"
the north star names as the target to beat for a11 + a12; the reference's Triton transform kernels are not part of it.

Prints ONE JSON object.  The REFERENCE arm of bench.py (--impl reference) runs this file in a subprocess (a fault
inside a reference kernel must not take the caller with it) and attaches the object to its line as
"reference_gpu_kernels"."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    try:
        import torch

        from oracle import ref_kernels as R
        if not R.available():
            raise RuntimeError("oracle/_ref/libflatquant_ref_kernels.so not built (needs /root/reference at build time)")
        R.load()
        torch.cuda.set_device(a.device)
        dev = torch.device(f"cuda:{a.device}")
        m = 2048
        shapes = [("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096)]
        g = torch.Generator(device=dev).manual_seed(0)
        acts = {k: torch.randint(0, 256, (m, k // 2), dtype=torch.uint8, device=dev, generator=g) for k in (4096, 14336)}
        s_row = (torch.rand(m, device=dev, generator=g) * 0.05 + 0.01).half()
        layers = []
        for _ in range(a.layers):
            layers.append([(acts[k], torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev, generator=g),
                            (torch.rand(n, device=dev, generator=g) * 0.01 + 0.005).half()) for _, k, n in shapes])

        def step():
            for layer in layers:
                for x, w, s_col in layer:
                    R.sym_dequant(R.matmul(x, w), s_row, s_col)

        for _ in range(max(1, a.warmup)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        ops = sum(2 * k * n for _, k, n in shapes) * m * a.layers
        out = {"what": "reference deploy/kernels: CUTLASS Sm80 s4 GEMM (matmul_host) + sym_dequant_host, built for sm_100a "
                       "from the reference's sources; GEMM + dequant only, 32 layers x 4 shapes at 2048 tokens",
               "ms_per_step": ms * 32 / a.layers, "tokens_per_s": m / (ms * 1e-3) * a.layers / 32,
               "tops_gemm": ops / (ms * 1e-3) / 1e12, "steps": a.steps, "layers": a.layers,
               "timing": "CUDA events on the legacy default stream (where the reference launches)"}
    except Exception as e:  # noqa: BLE001 -- the baseline leg reports, it never raises
        out = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
