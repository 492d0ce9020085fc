"""Writes flatquant_b200/deploy/functional/data/hadamard_tables.npz: the K x K Hadamard factors the reference hard-codes
for the non-power-of-two sizes (deploy/functional/online_trans.py:161-4083, Sloane's library: had12, had20, had28, had36,
had40, had52, had60, had108, had140, had156, had172), as DATA: one bit per entry (1 = +1), np.packbits row-major.

Runs in the build container only (needs /root/reference); the module's CUDA-only imports are stubbed.
    python tools/make_hadamard_tables.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flatquant_b200", "deploy", "functional", "data",
                   "hadamard_tables.npz")
ORDERS = (12, 20, 28, 36, 40, 52, 60, 108, 140, 156, 172)


def load_reference_module():
    for name in ("fast_hadamard_transform", "deploy", "deploy.kernels", "deploy.kernels.kron_matmul",
                 "deploy.kernels.block_matmul"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["deploy.kernels.kron_matmul"].kron_matmul = None
    sys.modules["deploy.kernels.block_matmul"].block_matmul = None
    spec = importlib.util.spec_from_file_location("ref_online_trans", os.path.join(REF, "deploy", "functional", "online_trans.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_module()
    out = {}
    for k in ORDERS:
        h = getattr(ref, f"get_had{k}")().numpy()
        assert h.shape == (k, k) and set(np.unique(h)) <= {-1.0, 1.0}
        assert np.array_equal(h @ h.T, k * np.eye(k)), k
        out[f"had{k}"] = np.packbits((h > 0).astype(np.uint8).reshape(-1))
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
