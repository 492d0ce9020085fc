"""Golden vectors for the non-power-of-two online Hadamard transform, from the reference's own Python:
flatquant/hadamard_utils.py:89-110 `matmul_hadU` (pure-torch butterflies + the hard-coded K x K table of get_hadK), the
CPU twin of deploy/functional/online_trans.py:144-151 `matmul_hadU_cuda`.  Build container only:

    python tests/golden/make_golden_hadamard_kn.py     ->  tests/golden/hadamard_kn.npz

Inputs are fp16-representable (the kernels take fp16); outputs are the reference's fp32 results."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
SIZES = (12 * 8, 20 * 16, 28 * 32, 40 * 8, 36 * 8, 52 * 8, 60 * 8, 108 * 8, 140 * 8, 156 * 8, 172 * 8, 172 * 64, 28 * 512)


def main():
    sys.path.insert(0, REF)
    sys.modules["fast_hadamard_transform"] = types.ModuleType("fast_hadamard_transform")
    sys.modules["flatquant.utils"] = types.ModuleType("flatquant.utils")  # needs accelerate otherwise
    sys.modules["flatquant.utils"].cleanup_memory = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location("ref_hadamard_utils", f"{REF}/flatquant/hadamard_utils.py")
    hu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hu)
    torch.manual_seed(31)
    out = {}
    for n in SIZES:
        rows = 2 if n > 4096 else 3
        x = torch.randn(rows, n).half().float()
        out[f"x_{n}"] = x.half().numpy()
        out[f"y_{n}"] = hu.matmul_hadU(x).numpy()
    np.savez_compressed(f"{OUT}/hadamard_kn.npz", **out)
    print("wrote hadamard_kn.npz", os.path.getsize(f"{OUT}/hadamard_kn.npz"))


if __name__ == "__main__":
    main()
