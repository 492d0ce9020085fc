"""ctypes binding of libflatq_b200.so (the C ABI declared in include/flatq_b200.h).

The library is the product: there is no Python/torch/CPU fallback.  If it is missing (or the
device is not sm_100) every op raises RuntimeError -- loudly, never silently.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLATQ_B200_LIB: load another build of the same library (A/B timing of kernel variants); never a fallback
LIB_PATH = os.environ.get("FLATQ_B200_LIB") or os.path.join(_HERE, "lib", "libflatq_b200.so")

# name -> (restype, argtypes); must list every symbol of include/flatq_b200.h
SIGNATURES = {
    "flatq_abi_version": (c_int, []),
    "flatq_last_error": (c_char_p, []),
    "flatq_device_supported": (c_int, []),
    "flatq_pack_i4": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "flatq_unpack_i4": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "flatq_sym_quant": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "flatq_sym_dequant": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int,
                                  c_void_p]),
    "flatq_gemm_i4_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "flatq_gemm_i4": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flatq_linear4bit_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "flatq_linear4bit": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                 c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "flatq_kron_transform": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "flatq_transform_quant": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "flatq_quantize_rows": (c_int, [c_void_p, c_int64, c_int64, c_float, c_float, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "flatq_act_quant": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int64, c_void_p, c_void_p,
                                c_void_p]),
    "flatq_gemm_w4a8_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "flatq_gemm_w4a8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                c_int, c_void_p, c_size_t, c_void_p]),
    "flatq_gemm_w4a8_silu_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "flatq_gemm_w4a8_silu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "flatq_linear_fused_silu_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "flatq_linear_fused_silu": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                                        c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flatq_gemm_w4a8_grouped_workspace_bytes": (c_size_t, [c_int]),
    "flatq_gemm_w4a8_grouped": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                        c_int64, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "flatq_hadamard": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "flatq_kv_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_int, c_int, c_void_p]),
    "flatq_kv_decode_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    # tensor parallel (section (e) of the header): flatq_tp_ctx* and the pointer arrays travel as void*
    "flatq_gemm_w4a8_rowparallel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "flatq_peer_reduce_bcast": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "flatq_peer_allreduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "flatq_add_rmsnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "flatq_hadamard_kn": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p]),
    "flatq_kv_quant": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_int, c_int, c_int, c_void_p]),
    "flatq_kv_decode_i4": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "flatq_kv_decode_i4_split_workspace_bytes": (c_size_t, [c_int, c_int]),
    "flatq_kv_decode_i4_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "flatq_rmsnorm": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "flatq_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "flatq_silu_mul": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "flatq_linear_fused_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "flatq_linear_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library; RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"flatquant_b200: CUDA library {LIB_PATH} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.flatq_abi_version() != 1:
            raise RuntimeError("flatquant_b200: ABI version mismatch between _lib.py and libflatq_b200.so")
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = load().flatq_last_error().decode(errors="replace")
        raise RuntimeError(f"flatq {what} failed ({status}): {msg}")
