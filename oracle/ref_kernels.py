"""TEST / BASELINE INFRASTRUCTURE: ctypes access to the reference's OWN native kernels (CUTLASS s4 GEMM, sym_quant,
sym_dequant, FlashInfer INT4 decode / KV append), compiled from the sources under /root/reference by
`__graft_entry__.build()` into oracle/_ref/libflatquant_ref_kernels.so (see oracle/ref_kernels_shim.cu).

Only tests/ and the baseline leg of the benchmarks may import this module; nothing under flatquant_b200/ does.  The
wrappers take CUDA torch tensors, allocate outputs like deploy/kernels/bindings.cpp does, and launch on the legacy
default stream (as the reference does); callers synchronise."""
import ctypes
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libflatquant_ref_kernels.so")
_lib = None


class RefKernelError(RuntimeError):
    pass


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        if not available():
            raise RefKernelError(f"{LIB_PATH} has not been built (needs /root/reference at build time)")
        lib = ctypes.CDLL(LIB_PATH)
        vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
        lib.ref_last_error.restype = ctypes.c_char_p
        lib.ref_matmul.argtypes = [vp, vp, u32, u32, u32, vp]
        lib.ref_sym_quant.argtypes = [vp, vp, u32, u32, u32, vp]
        lib.ref_sym_dequant.argtypes = [vp, vp, vp, u32, u32, vp]
        lib.ref_batch_decode_i4.argtypes = [vp] * 7 + [i32] * 5
        lib.ref_init_kv_i4.argtypes = [vp] * 10 + [i32] * 5
        lib.ref_append_kv_i4.argtypes = [vp] * 9 + [i32] * 5
        _lib = lib
    return _lib


def _ck(rc, what):
    if rc:
        raise RefKernelError(f"reference {what} failed ({rc}): {load().ref_last_error().decode(errors='replace')}")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def matmul(a, b):
    """deploy._CUDA.matmul (bindings.cpp:9-25): packed A [M, K/2] x packed B [N, K/2]^T -> int32 [M, N]."""
    import torch
    assert a.is_cuda and b.is_cuda and a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype == torch.uint8
    m, n, k = a.shape[0], b.shape[0], a.shape[1] * 2
    c = torch.empty((m, n), dtype=torch.int32, device=a.device)
    _ck(load().ref_matmul(_p(a), _p(b), m, n, k, _p(c)), "matmul_host")
    return c


def sym_quant(x, scale):
    """deploy._CUDA.sym_quant (bindings.cpp:27-44): x f16 [rows, cols], scale f16 [rows] -> packed [rows, cols/2]."""
    import torch
    assert x.is_cuda and x.dtype == scale.dtype == torch.float16 and x.is_contiguous() and scale.is_contiguous()
    rows, cols = x.shape
    q = torch.empty((rows, cols // 2), dtype=torch.uint8, device=x.device)
    _ck(load().ref_sym_quant(_p(x), _p(scale), rows, cols, cols // 2, _p(q)), "sym_quant_host")
    return q


def sym_dequant(q, scale_row, scale_col):
    """deploy._CUDA.sym_dequant (bindings.cpp:47-87, bits = 32): int32 [rows, cols] -> f16."""
    import torch
    assert q.is_cuda and q.dtype == torch.int32 and scale_row.dtype == scale_col.dtype == torch.float16
    rows, cols = q.shape
    x = torch.empty((rows, cols), dtype=torch.float16, device=q.device)
    _ck(load().ref_sym_dequant(_p(q), _p(scale_row), _p(scale_col), rows, cols, _p(x)), "sym_dequant_host")
    return x


def batch_decode_i4(q, kv_data, kv_param, kv_indptr, kv_indices, last_page_offset, layer_idx):
    """deploy._CUDA.batch_decode_i4 (bindings.cpp:118-170): q f16 [B, N, 128] over the paged INT4 cache
    kv_data [pages, L, 2, N, P, 64] u8 / kv_param [pages, L, 2, N, P, 2] f16 -> o f16 [B, N, 128]."""
    import torch
    assert q.dtype == torch.float16 and kv_data.dtype == torch.uint8 and kv_param.dtype == torch.float16
    o = torch.empty_like(q)
    _ck(load().ref_batch_decode_i4(_p(o), _p(q), _p(kv_data), _p(kv_param), _p(kv_indptr), _p(kv_indices),
                                   _p(last_page_offset), kv_data.shape[1], layer_idx, kv_data.shape[3], kv_data.shape[4],
                                   q.shape[0]), "FlashInferBatchDecodeKernel_i4")
    return o
