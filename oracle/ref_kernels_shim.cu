// TEST / BASELINE INFRASTRUCTURE -- not part of the product library.
//
// C-ABI wrappers around the reference's own native kernels, compiled from the sources WHERE THEY LIE under
// /root/reference (deploy/kernels/{gemm.cu, quant.cu, flashinfer.cu} + the vendored third-party/cutlass headers) by
// __graft_entry__.build() into oracle/_ref/libflatquant_ref_kernels.so (git-ignored; travels to the GPU box with the
// snapshot).  No reference source is copied into this repository: this file only declares and calls
//     matmul_host        deploy/kernels/include/gemm.h:6-13   (CUTLASS Sm80 s4 x s4 -> s32 GEMM, gemm.cu:7-46)
//     sym_quant_host     deploy/kernels/include/quant.h:6-13  (quant.cu:13-63)
//     sym_dequant_host   deploy/kernels/include/quant.h:16-23 (quant.cu:66-101)
//     FlashInfer{BatchDecode,InitKv,AppendKv}Kernel_i4<128>   deploy/kernels/include/flashinfer.h:6-31 (flashinfer.cu)
// exactly as deploy/kernels/bindings.cpp:9-245 does, minus torch.  The reference's `ensure()` throws
// std::runtime_error (include/common.h:16-23); an exception crossing a ctypes call would abort the process, so every
// wrapper catches and returns a status.  All reference kernels launch on the legacy default stream.
//
// Used by tests/test_zy_reference_cuda_kernels_gpu.py (bit-level GPU parity of flatq_gemm_i4 / flatq_sym_quant /
// flatq_sym_dequant(mode 1) / flatq_kv_decode_i4 against the reference's kernels on the same B200) and by
// tests/tools/ref_kernels_bench.py (the reference's GEMM + dequant timed by the reference arm of bench.py).
#include <flashinfer.h>
#include <gemm.h>
#include <quant.h>

#include <string>

namespace {
thread_local std::string g_ref_err;

template <typename F>
int guarded(F&& f) {
    try {
        f();
    } catch (const std::exception& e) {
        g_ref_err = e.what();
        return -1;
    } catch (...) {
        g_ref_err = "unknown C++ exception";
        return -1;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_ref_err = cudaGetErrorString(e);
        return -2;
    }
    return 0;
}
}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_ref_err.c_str(); }

int ref_matmul(const uint8_t* A, const uint8_t* B, uint32_t M, uint32_t N, uint32_t K, int32_t* C) {
    return guarded([&] { matmul_host(A, B, M, N, K, C); });
}

int ref_sym_quant(const void* x, const void* scale, uint32_t rows, uint32_t cols_src, uint32_t cols_dst, uint8_t* q) {
    return guarded([&] {
        sym_quant_host(reinterpret_cast<const half*>(x), reinterpret_cast<const half*>(scale), rows, cols_src, cols_dst, q);
    });
}

int ref_sym_dequant(const int32_t* q, const void* scale_row, const void* scale_col, uint32_t rows, uint32_t cols, void* x) {
    return guarded([&] {
        sym_dequant_host(q, reinterpret_cast<const half*>(scale_row), reinterpret_cast<const half*>(scale_col), rows, cols,
                         reinterpret_cast<half*>(x));
    });
}

int ref_batch_decode_i4(void* o, void* q, void* kv_data, void* kv_param, int32_t* kv_indptr, int32_t* kv_indices,
                        int32_t* last_page_offset, int num_layers, int layer_idx, int num_heads, int page_size,
                        int batch_size) {
    return guarded([&] {
        FlashInferBatchDecodeKernel_i4<128>(reinterpret_cast<nv_half*>(o), reinterpret_cast<nv_half*>(q), kv_data,
                                            reinterpret_cast<nv_half2*>(kv_param), kv_indptr, kv_indices, last_page_offset,
                                            num_layers, layer_idx, num_heads, page_size, batch_size);
    });
}

int ref_init_kv_i4(void* kv_data, void* kv_param, int32_t* kv_indptr, int32_t* kv_indices, int32_t* last_page_offset,
                   void* key, void* value, void* key_param, void* value_param, int32_t* seqlen_indptr, int num_layers,
                   int layer_idx, int num_heads, int page_size, int batch_size) {
    return guarded([&] {
        FlashInferInitKvKernel_i4<128>(kv_data, reinterpret_cast<nv_half2*>(kv_param), kv_indptr, kv_indices,
                                       last_page_offset, key, value, reinterpret_cast<nv_half2*>(key_param),
                                       reinterpret_cast<nv_half2*>(value_param), seqlen_indptr, num_layers, layer_idx,
                                       num_heads, page_size, batch_size);
    });
}

int ref_append_kv_i4(void* kv_data, void* kv_param, int32_t* kv_indptr, int32_t* kv_indices, int32_t* last_page_offset,
                     void* key, void* value, void* key_param, void* value_param, int num_layers, int layer_idx,
                     int num_heads, int page_size, int batch_size) {
    return guarded([&] {
        FlashInferAppendKvKernel_i4<128>(kv_data, reinterpret_cast<nv_half2*>(kv_param), kv_indptr, kv_indices,
                                         last_page_offset, key, value, reinterpret_cast<nv_half2*>(key_param),
                                         reinterpret_cast<nv_half2*>(value_param), num_layers, layer_idx, num_heads,
                                         page_size, batch_size);
    });
}

}  // extern "C"
