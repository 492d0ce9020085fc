// C-ABI entry points of libflatq_b200.so (see include/flatq_b200.h for the contract and the
// reference interfaces each symbol replaces).  Host glue only: argument checks, TMA descriptors,
// launch configuration.  No torch, no allocation.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <string>

#include "../../include/flatq_b200.h"
#include <atomic>
#include <random>
#include "elementwise.cuh"
#include "gemm_w4a8.cuh"
#include "gemm_w4a8_2sm.cuh"
#include "gemv_w4a8.cuh"
#include "gemm_w4a8_skinny.cuh"
#include "flatlinear_decode.cuh"
#include "glue.cuh"
#include "kron_quant.cuh"
#include "kron_quant_tc.cuh"
#include "kron_quant_ts.cuh"
#include "qattention.cuh"
#include "decode_attn_mma.cuh"
#include "kv_fused.cuh"
#include "peer_reduce.cuh"

using namespace flatq;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define FLATQ_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) return fail(FLATQ_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define FLATQ_CHECK(cond, code, ...) \
    do {                             \
        if (!(cond)) return fail(code, __VA_ARGS__); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cudaFuncSetAttribute is per device: a process that drives several GPUs must opt every kernel in on each of them.
// One flag (or the largest size opted in so far) per device ordinal; races are benign -- the call is idempotent.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDevices ? d : -1;
}
struct PerDeviceOnce {
    bool done[kMaxDevices] = {};
    bool first_time() {   // true once per device (always true when the device cannot be determined)
        const int d = current_device();
        if (d < 0) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
struct PerDeviceMax {
    size_t seen[kMaxDevices] = {};
    bool grows(size_t v) {  // true when v exceeds what this device has been opted in for
        const int d = current_device();
        if (d < 0) return true;
        if (v <= seen[d]) return false;
        seen[d] = v;
        return true;
    }
};

int num_sms() {
    static int cached[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int grid_for(size_t n_items, int block, int max_waves = 8) {
    size_t g = (n_items + block - 1) / block;
    size_t cap = (size_t)num_sms() * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- launches with programmatic dependent launch (PDL): the kernel's prologue overlaps the previous kernel's tail
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool use_pdl = getenv("FLATQ_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- TMA descriptors through the runtime-resolved driver entry point (no link-time libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2-D byte tensor [rows, cols_bytes] with row pitch `pitch`; box = [box_rows, box_cols] bytes.
int make_map_u8(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols_bytes, uint64_t pitch,
                uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    FLATQ_CHECK(fn != nullptr, FLATQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    FLATQ_CHECK(aligned16(ptr) && pitch % 16 == 0, FLATQ_ERR_INVALID, "TMA operand must be 16-byte aligned (pitch %llu)",
                (unsigned long long)pitch);
    cuuint64_t dims[2] = {cols_bytes, rows};
    cuuint64_t strides[1] = {pitch};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLATQ_CHECK(r == CUDA_SUCCESS, FLATQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return FLATQ_OK;
}

// ---- GEMM launch
template <int TN, int MODE>
int launch_gemm_t(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_w4a8_kernel<TN, MODE>;
    constexpr int smem = GemmSmem<TN>::kBytes;
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int tiles = ((p.N + kGemmBlockW - 1) / kGemmBlockW) * ((p.M + TN - 1) / TN) * (p.k_splits > 1 ? p.k_splits : 1);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    FLATQ_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, st, tw, tx, p));
    return FLATQ_OK;
}

// CTA-pair kernel (prefill): grid = 2 x min(tiles of 256 weight rows x 256 tokens, SM pairs)
template <int MODE>
int launch_gemm_2sm(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_w4a8_2sm_kernel<MODE>;
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k2smBytes));
    const int nb_w2 = ((p.N + kGemmBlockW - 1) / kGemmBlockW + 1) / 2, nb_m = (p.M + k2smTokens - 1) / k2smTokens;
    const int tiles = nb_w2 * nb_m, pairs = num_sms() / 2;
    const int grid = 2 * (tiles < pairs ? tiles : pairs);
    FLATQ_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), k2smBytes, st, tw, tx, p));
    return FLATQ_OK;
}

template <int MODE>
int launch_gemm_m(int tn, const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
    switch (tn) {
        case 16: return launch_gemm_t<16, MODE>(tw, tx, p, st);
        case 32: return launch_gemm_t<32, MODE>(tw, tx, p, st);
        case 64: return launch_gemm_t<64, MODE>(tw, tx, p, st);
        case 128: return launch_gemm_t<128, MODE>(tw, tx, p, st);
        default: return launch_gemm_t<256, MODE>(tw, tx, p, st);
    }
}

// w_packed [N, K/2]; x8 [M, K] int8 in the permuted operand format
// Kernel-variant switches (A/B timing and the parity tests of the non-default variants; defaults from the environment):
//   FLATQ_GEMM_1SM=1  prefill on the single-CTA kernel instead of CTA pairs
//   FLATQ_STREAMK=1   stream-K on the partial last wave of the single-CTA kernel (measured: no gain, see
//                     profiles/r01_streamk_power_bound.md)
static int g_gemm_1sm = -1, g_streamk = -1;
static bool opt_gemm_1sm() {
    if (g_gemm_1sm < 0) g_gemm_1sm = getenv("FLATQ_GEMM_1SM") != nullptr;
    return g_gemm_1sm != 0;
}
static bool opt_streamk() {
    if (g_streamk < 0) g_streamk = getenv("FLATQ_STREAMK") != nullptr;
    return g_streamk != 0;
}

// stream-K workspace of the prefill GEMM: per CTA one 128 x TN int32 partial + 16 flags (see GemmParams::streamk)
static size_t streamk_ws_bytes(int tn) {
    return align_up((size_t)num_sms() * kGemmBlockW * tn * 4, 256) + align_up((size_t)num_sms() * 16 * 8, 256);
}

static unsigned long long next_epoch() {
    static std::atomic<unsigned long long> ctr{0};
    static const unsigned long long salt = [] {
        std::random_device rd;
        return ((unsigned long long)rd() << 32) ^ ((unsigned long long)rd() << 11);
    }();
    unsigned long long e = salt + ctr.fetch_add(1) + 1;
    return e ? e : 1;
}

int launch_gemm(const uint8_t* w_packed, const void* x8, int64_t M, int64_t N, int64_t K, int acc_shift, void* out,
                int mode, const float* sx, const float* sw, const void* bias, cudaStream_t st, int k_splits = 1,
                void* sk_ws = nullptr, size_t sk_ws_bytes = 0, bool silu_pairs = false) {
    FLATQ_CHECK(M > 0 && N > 0 && K > 0, FLATQ_ERR_INVALID, "empty GEMM %lldx%lldx%lld", (long long)M, (long long)N,
                (long long)K);
    FLATQ_CHECK(K % 32 == 0, FLATQ_ERR_INVALID, "K=%lld must be a multiple of 32", (long long)K);
    FLATQ_CHECK(K <= 65536 && M < (1ll << 31) && N < (1ll << 31), FLATQ_ERR_UNSUPPORTED, "GEMM too large");
    FLATQ_CHECK(mode == kEpiI32 || N % 2 == 0, FLATQ_ERR_UNSUPPORTED, "N=%lld must be even for 16-bit outputs", (long long)N);
    FLATQ_CHECK(mode == kEpiI32 || (M + 256) * N < (1ll << 32), FLATQ_ERR_UNSUPPORTED,
                "output of %lld x %lld exceeds the 32-bit element index of the store path", (long long)M, (long long)N);
    const int tn = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
    // prefill: two SMs share a 256 x 256 tile (each loads half of the activation tile)
    const bool pair = tn == 256 && k_splits == 1 && !opt_gemm_1sm();
    CUtensorMap tw, tx;
    int rc = make_map_u8(&tw, w_packed, N, K / 2, K / 2, kGemmBlockK / 2, kGemmBlockW, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_u8(&tx, x8, M, K, K, 128, pair ? 128 : tn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.acc_shift = acc_shift;
    p.out = out; p.s_x = sx; p.s_w = sw; p.bias = bias; p.prof = nullptr; p.dbg = 0; p.k_splits = k_splits;
    p.streamk = 0; p.sk_part = nullptr; p.sk_flags = nullptr; p.epoch = 0;
    p.grp_off = nullptr; p.grp_tile = nullptr; p.n_groups = 0; p.t_upper = 0; p.rows_per_owner = 0; p.tp_rank = 0; p.silu_pairs = 0;
    {
        // Stream-K on the partial last wave: worth it when the tail's K stages, dealt to all SMs, finish at least two
        // stages earlier than whole tiles would, and every part is long enough to amortise its fix-up.
        const int64_t G = num_sms(), T = ((N + kGemmBlockW - 1) / kGemmBlockW) * ((M + tn - 1) / tn);
        const int64_t KB = (K + kGemmBlockK - 1) / kGemmBlockK, R = T % G;
        const int min_share = 1;
        if (opt_streamk() && !pair && mode != kEpiI32 && k_splits == 1 && sk_ws && aligned16(sk_ws) && sk_ws_bytes >= streamk_ws_bytes(tn) &&
            T > G && R > 0 && R * KB / G >= min_share && KB - (R * KB + G - 1) / G >= 2) {
            p.streamk = 1;
            p.sk_part = reinterpret_cast<int32_t*>(sk_ws);
            p.sk_flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(sk_ws) +
                                                               align_up((size_t)G * kGemmBlockW * tn * 4, 256));
            p.epoch = next_epoch();
        }
    }
    FLATQ_CHECK(!silu_pairs || (pair && mode == kEpiF16), FLATQ_ERR_UNSUPPORTED, "SiLU epilogue outside the CTA-pair fp16 kernel");
    p.silu_pairs = silu_pairs ? 1 : 0;
    if (pair) {
        switch (mode) {
            case kEpiI32: return launch_gemm_2sm<kEpiI32>(tw, tx, p, st);
            case kEpiF16: return launch_gemm_2sm<kEpiF16>(tw, tx, p, st);
            case kEpiBF16: return launch_gemm_2sm<kEpiBF16>(tw, tx, p, st);
        }
    }
    switch (mode) {
        case kEpiI32: return launch_gemm_m<kEpiI32>(tn, tw, tx, p, st);
        case kEpiF16: return launch_gemm_m<kEpiF16>(tn, tw, tx, p, st);
        case kEpiBF16: return launch_gemm_m<kEpiBF16>(tn, tw, tx, p, st);
    }
    return fail(FLATQ_ERR_INVALID, "bad epilogue mode %d", mode);
}

// ---- grouped (mixture-of-experts) GEMM: the single-CTA kernel with the schedule of csrc/grouped_sched.h
template <int TN, int MODE>
int launch_gemm_grouped_t(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_w4a8_kernel<TN, MODE, false, true>;
    constexpr int smem = GemmSmem<TN>::kBytes;
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = p.t_upper < num_sms() ? p.t_upper : num_sms();
    FLATQ_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, st, tw, tx, p));
    return FLATQ_OK;
}

template <int MODE>
int launch_gemm_grouped_m(int tn, const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, cudaStream_t st) {
    switch (tn) {
        case 16: return launch_gemm_grouped_t<16, MODE>(tw, tx, p, st);
        case 32: return launch_gemm_grouped_t<32, MODE>(tw, tx, p, st);
        case 64: return launch_gemm_grouped_t<64, MODE>(tw, tx, p, st);
        case 128: return launch_gemm_grouped_t<128, MODE>(tw, tx, p, st);
        default: return launch_gemm_grouped_t<256, MODE>(tw, tx, p, st);
    }
}

// ---- Kronecker transform (+quant) launch
template <int KT2>
int launch_kron_t(const KronParams& p, int threads, size_t smem, int grid, cudaStream_t st) {
    auto kern = kron_quant_mma_kernel<KT2>;
    static PerDeviceMax attr;
    if (smem > 48 * 1024 && attr.grows(smem))
        FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLATQ_CUDA(launch_pdl(kern, dim3(grid), dim3(threads), smem, st, p));
    return FLATQ_OK;
}

int launch_kron_tc(const KronParams& p, cudaStream_t st);

int launch_kron(KronParams p, cudaStream_t st) {
    FLATQ_CHECK(p.tokens > 0, FLATQ_ERR_INVALID, "no tokens");
    FLATQ_CHECK(aligned16(p.x), FLATQ_ERR_INVALID, "x must be 16-byte aligned");
    {
        static const bool no_tc = getenv("FLATQ_KRON_HMMA") != nullptr;  // debugging aid: force the HMMA kernel
        if (!no_tc) {
            const int rc = launch_kron_tc(p, st);
            if (rc != 1) return rc;
        }
    }
    FLATQ_CHECK(p.n1 >= 1 && p.n2 >= 16 && p.n1 <= 256 && p.n2 <= 256 && p.n2 % 16 == 0, FLATQ_ERR_UNSUPPORTED,
                "unsupported Kronecker factor sizes (%d, %d): need n2 %% 16 == 0 and n1, n2 <= 256", p.n1, p.n2);
    FLATQ_CHECK(aligned16(p.x), FLATQ_ERR_INVALID, "x must be 16-byte aligned");
    p.p1 = (p.n1 + 15) / 16 * 16;
    const int wpt = p.p1 / 16;
    const int kt2 = p.n2 / 16;
    const int max_threads = kt2 <= 8 ? 512 : 256;
    FLATQ_CHECK(wpt * 32 <= max_threads, FLATQ_ERR_UNSUPPORTED, "Kronecker factors (%d, %d) need %d warps per token",
                p.n1, p.n2, wpt);
    int tpc = 8 / wpt;
    if (tpc < 1) tpc = 1;
    while (tpc > 1 && (tpc * wpt * 32 > max_threads ||
                       kron_smem_bytes(p.n1, p.n2, p.p1, tpc, p.left != nullptr, p.right != nullptr) > 100 * 1024))
        --tpc;
    p.tokens_per_cta = tpc;
    const size_t smem = kron_smem_bytes(p.n1, p.n2, p.p1, tpc, p.left != nullptr, p.right != nullptr);
    FLATQ_CHECK(smem <= 227 * 1024, FLATQ_ERR_UNSUPPORTED, "Kronecker factors (%d, %d) exceed shared memory", p.n1, p.n2);
    const int threads = tpc * wpt * 32;
    int64_t ctas = (p.tokens + tpc - 1) / tpc;
    const int64_t per_sm = (int64_t)(220 * 1024 / (smem + 1024));
    const int64_t cap = (int64_t)num_sms() * (per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm);
    const int grid = (int)(ctas < cap ? ctas : cap);
    switch (kt2) {
#define FLATQ_KRON_CASE(n) case n: return launch_kron_t<n>(p, threads, smem, grid, st);
        FLATQ_KRON_CASE(1) FLATQ_KRON_CASE(2) FLATQ_KRON_CASE(3) FLATQ_KRON_CASE(4) FLATQ_KRON_CASE(5)
        FLATQ_KRON_CASE(6) FLATQ_KRON_CASE(7) FLATQ_KRON_CASE(8) FLATQ_KRON_CASE(9) FLATQ_KRON_CASE(10)
        FLATQ_KRON_CASE(11) FLATQ_KRON_CASE(12) FLATQ_KRON_CASE(13) FLATQ_KRON_CASE(14) FLATQ_KRON_CASE(15)
        FLATQ_KRON_CASE(16)
#undef FLATQ_KRON_CASE
    }
    return fail(FLATQ_ERR_UNSUPPORTED, "n2=%d", p.n2);
}


// ---- tcgen05 Kronecker transform + quantise (the shapes of the Llama linears); other shapes use the HMMA kernel
template <int N2, int N1P, int T, bool OUT8, bool EXACT>
int launch_kron_tc_t(const KronParams& p, cudaStream_t st) {
    using C = KronTc<N2, N1P, T>;
    auto kern = kron_quant_tc_kernel<N2, N1P, T, OUT8, EXACT>;
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    CUtensorMap tx;  // x as a 2-D byte tensor [tokens * n1 rows][n2 * 2 bytes]
    int rc = make_map_u8(&tx, p.x, (uint64_t)p.tokens * p.n1, (uint64_t)N2 * 2, (uint64_t)N2 * 2, 128, C::kXRows,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    CUtensorMap tr = tx, tl = tx;  // EXACT: the factors as row-major 2-D byte tensors, 64-column (128 B) panels
    if (EXACT) {
        rc = make_map_u8(&tr, p.right, N2, (uint64_t)N2 * 2, (uint64_t)N2 * 2, 128, N2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = make_map_u8(&tl, p.left, N1P, (uint64_t)N1P * 2, (uint64_t)N1P * 2, 128, N1P, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int groups = (p.tokens + T - 1) / T;
    const int cap = C::kMinCtas * num_sms();  // persistent: 2-8x more CTAs (one group each) measured 15-45 % slower
    FLATQ_CUDA(launch_pdl(kern, dim3(groups < cap ? groups : cap), dim3(C::kThreads), (size_t)C::kSmemBytes, st, tx, tr, tl, p));
    return FLATQ_OK;
}

template <int N2, int N1P, int T>
int launch_kron_tc_o(const KronParams& p, cudaStream_t st) {
    // EXACT: no padding and both factors present and TMA-addressable -> factor panels by TMA, MN-major operands
    if (p.n1 == N1P && p.left && p.right && aligned16(p.left) && aligned16(p.right)) return p.xq8 ? launch_kron_tc_t<N2, N1P, T, true, true>(p, st) : launch_kron_tc_t<N2, N1P, T, false, true>(p, st);
    return p.xq8 ? launch_kron_tc_t<N2, N1P, T, true, false>(p, st) : launch_kron_tc_t<N2, N1P, T, false, false>(p, st);
}

// ---- second tcgen05 design (kron_quant_ts.cuh): intermediate in TMEM, several token groups in flight per CTA
template <int N1, int N2, int T, bool OUT8>
int launch_kron_ts_t(const KronParams& p, cudaStream_t st) {
    using C = KronTs<N1, N2, T>;
    auto kern = kron_quant_ts_kernel<N1, N2, T, OUT8>;
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    CUtensorMap tx, tr;  // x as a 2-D byte tensor [tokens * n1 rows][n2 * 2 bytes]; R row-major, 64-column panels
    int rc = make_map_u8(&tx, p.x, (uint64_t)p.tokens * N1, (uint64_t)N2 * 2, (uint64_t)N2 * 2, 128, C::kK1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_u8(&tr, p.right, N2, (uint64_t)N2 * 2, (uint64_t)N2 * 2, 128, N2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const int groups = (p.tokens + T - 1) / T;
    const int grid = groups < num_sms() ? groups : num_sms();
    FLATQ_CUDA(launch_pdl(kern, dim3(grid), dim3(C::kThreads), (size_t)C::kSmemBytes, st, tx, tr, p));
    return FLATQ_OK;
}

template <int N1, int N2, int T>
int launch_kron_ts_o(const KronParams& p, cudaStream_t st) {
    return p.xq8 ? launch_kron_ts_t<N1, N2, T, true>(p, st) : launch_kron_ts_t<N1, N2, T, false>(p, st);
}

// 128 x 128 fp16 identity for left-only requests (the o_proj transform over the heads axis, Z = O^T . A with A [heads, 128]):
// they run on the two-factor kernel with R = I, which is exact (every product is x * 1 or x * 0) and three times faster
// than the legacy-MMA kernel that served them in round 1.
__device__ __half g_eye128[128 * 128];
__global__ void eye128_init_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 128 * 128) g_eye128[i] = __float2half((i >> 7) == (i & 127) ? 1.f : 0.f);
}
static const __half* eye128(cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.first_time()) eye128_init_kernel<<<64, 256, 0, st>>>();   // same stream as its first consumer
    void* ptr = nullptr;
    if (cudaGetSymbolAddress(&ptr, g_eye128) != cudaSuccess) return nullptr;
    return reinterpret_cast<const __half*>(ptr);
}

// returns 1 if the request is not one of the instantiated (n1, n2) pairs
int launch_kron_ts(const KronParams& p_in, cudaStream_t st) {
    KronParams p = p_in;
    if (p.left && !p.right && p.n2 == 128 && (p.n1 == 64 || p.n1 == 32 || p.n1 == 16)) p.right = eye128(st);
    if (!p.left || !p.right || !aligned16(p.left) || !aligned16(p.right)) return 1;
    if (p.n2 == 64) {
        if (p.n1 == 64) return launch_kron_ts_o<64, 64, 2>(p, st);     // 4096 (Llama-3-8B hidden)
        if (p.n1 == 56) return launch_kron_ts_o<56, 64, 2>(p, st);     // 3584 (70B down_proj at tp 8)
        if (p.n1 == 32) return launch_kron_ts_o<32, 64, 4>(p, st);     // 2048
    } else if (p.n2 == 128) {
        if (p.n1 == 64) return launch_kron_ts_o<64, 128, 2>(p, st);    // 8192 (70B hidden; left-only: 64 heads)
        if (p.n1 == 32) return launch_kron_ts_o<32, 128, 4>(p, st);    // left-only: o_proj over 32 heads x 128 (Llama-3-8B)
        if (p.n1 == 16) return launch_kron_ts_o<16, 128, 4>(p, st);    // left-only: 16 heads per rank (tp 2 / 70B tp 4)
        if (p.n1 == 112) return launch_kron_ts_o<112, 128, 1>(p, st);  // 14336 (Llama-3-8B intermediate)
        if (p.n1 == 128) return launch_kron_ts_o<128, 128, 1>(p, st);  // 16384 (DeepSeek-V3 wo)
    }
    return 1;
}

// returns 1 if no tcgen05 instantiation covers the request (caller falls back to the HMMA kernel)
int launch_kron_tc(const KronParams& p, cudaStream_t st) {
    if (!p.do_quant || p.z_out || p.tokens <= 0 || (p.n1 * p.n2) % 32 != 0) return 1;
    if ((p.xq8 != nullptr) == (p.xq4 != nullptr)) return 1;  // exactly one code format per launch
    {
        static const bool v1 = getenv("FLATQ_KRON_TC1") != nullptr;   // A/B aid: force the first tcgen05 design
        if (!v1) {
            const int rc = launch_kron_ts(p, st);
            if (rc != 1) return rc;
        }
    }
    if (p.n2 == 64) {
        if (p.n1 == 64) return launch_kron_tc_o<64, 64, 2>(p, st);
        if (p.n1 > 32 && p.n1 < 64) return launch_kron_tc_o<64, 64, 1>(p, st);
        if (p.n1 == 32) return launch_kron_tc_o<64, 32, 2>(p, st);
    } else if (p.n2 == 128) {
        if (p.n1 > 32 && p.n1 <= 64) return launch_kron_tc_o<128, 64, 1>(p, st);
        if (p.n1 > 64 && p.n1 <= 96) return launch_kron_tc_o<128, 96, 1>(p, st);
        if (p.n1 > 96 && p.n1 <= 112) return launch_kron_tc_o<128, 112, 1>(p, st);
        if (p.n1 > 112 && p.n1 <= 128) return launch_kron_tc_o<128, 128, 1>(p, st);
    }
    return 1;
}

// ---- one-launch decode FlatLinear (flatlinear_decode.cuh): transform + quantise in the GEMV's own prologue
static int8_t* g_decode_dump_codes = nullptr;   // tests only (flatq_debug_set_decode_dump)
static float* g_decode_dump_sx = nullptr;

static bool flat_decode_ok(int64_t M, int64_t N, int n1, int n2, const void* x, const void* left, const void* right,
                           const void* w, bool silu) {
    static const bool off = getenv("FLATQ_NO_FUSED_DECODE") != nullptr;   // A/B aid: transform kernel + GEMV
    if (off || M < 1 || M > 2 || (!left && !right) || n1 % 16 || n1 < 16 || n1 > 128 || (n2 != 64 && n2 != 128)) return false;
    const int64_t K = (int64_t)n1 * n2;
    if (K % 1024 || !aligned16(x) || !aligned16(w) || (left && !aligned16(left)) || (right && !aligned16(right))) return false;
    if (silu && N % 2) return false;
    const int cpl = (int)(K / 1024);
    if (n2 == 64) return cpl == 2 || cpl == 4;
    // K = 14336 (112 x 128, Llama-3-8B down_proj) stays on the transform kernel + GEMV: 6.9 MFLOP of transform per CTA and
    // 173 registers made the one-launch form slower there (16.7 vs 12.3 us per fused linear from a graph)
    return silu ? cpl == 8 : (cpl == 4 || cpl == 8);
}

template <int MT, int N2, int CPL, int R, int MINB, bool SILU>
cudaError_t launch_flat_decode_t(const FlatDecodeParams& q, cudaStream_t st) {
    auto kern = flatlinear_decode_kernel<MT, N2, CPL, R, MINB, SILU>;
    static PerDeviceOnce attr;
    if (attr.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    const size_t smem = flat_decode_smem_bytes(q.n1, N2, MT);
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGemvWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t units = SILU ? q.g.N / 2 : q.g.N, ctas = (units + kGemvWarps - 1) / kGemvWarps, cap = (int64_t)per_sm * num_sms();
    return launch_pdl(kern, dim3((unsigned)(cap < ctas ? cap : ctas)), dim3(kGemvWarps * 32), smem, st, q);
}

template <bool SILU>
cudaError_t launch_flat_decode(const FlatDecodeParams& q, cudaStream_t st) {
    const int cpl = q.g.K / 1024, key = (q.n2 == 64 ? 0 : 1000) + cpl * 10 + q.g.M;
    constexpr int R = SILU ? 2 : 1;
    switch (key) {
#define FLATQ_FD(N2, CPL, MINB)                                                                                \
    case (N2 == 64 ? 0 : 1000) + CPL * 10 + 1: return launch_flat_decode_t<1, N2, CPL, R, MINB, SILU>(q, st);  \
    case (N2 == 64 ? 0 : 1000) + CPL * 10 + 2: return launch_flat_decode_t<2, N2, CPL, R, MINB, SILU>(q, st);
        FLATQ_FD(64, 2, 3)      // K = 2048  (32 x 64)
        FLATQ_FD(64, 4, 2)      // K = 4096  (64 x 64): Llama-3-8B hidden
        FLATQ_FD(128, 8, 1)     // K = 8192  (64 x 128): Llama-3-70B hidden
#undef FLATQ_FD
        default: break;
    }
    if (!SILU) {
        switch (key) {
#define FLATQ_FD(N2, CPL, MINB)                                                                                 \
    case 1000 + CPL * 10 + 1: return launch_flat_decode_t<1, N2, CPL, 1, MINB, false>(q, st);                   \
    case 1000 + CPL * 10 + 2: return launch_flat_decode_t<2, N2, CPL, 1, MINB, false>(q, st);
            FLATQ_FD(128, 4, 2)     // K = 4096  (32 x 128): o_proj over 32 heads, left factor only
#undef FLATQ_FD
            default: break;
        }
    }
    return cudaErrorInvalidValue;
}

// decode batches (3..32 tokens): the weight-streaming mma.sync kernel of gemm_w4a8_skinny.cuh
static bool skinny_ok(int64_t M, int64_t N, int64_t K, const void* x, const void* w) {
    static const bool off = getenv("FLATQ_NO_SKINNY") != nullptr;   // A/B aid: back to the tcgen05 split-K route
    return !off && M >= 1 && M <= 32 && K % 128 == 0 && N >= 1 && aligned16(x) && aligned16(w);
}

template <int NB, bool SILU, bool XS>
cudaError_t launch_skinny_nb(const GemvParams& g, cudaStream_t st) {
    auto kern = gemm_w4a8_skinny_kernel<NB, SILU, XS>;
    const size_t smem = (size_t)skinny_red_bytes<NB>() + (XS ? (size_t)NB * 8 * (g.K + 16) : 0);
    static PerDeviceOnce attr;
    if (attr.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;   // resident CTAs per SM (registers and this call's shared memory decide): the grid is one wave
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSkinnyWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (g.N + kSkinnyRows - 1) / kSkinnyRows, cap = (int64_t)per_sm * num_sms();
    return launch_pdl(kern, dim3((unsigned)(tiles < cap ? tiles : cap)), dim3(kSkinnyWarps * 32), smem, st, g);
}

template <int NB, bool SILU>
cudaError_t launch_skinny_xs(const GemvParams& g, cudaStream_t st) {
    // staging the codes costs every CTA one pass over them: worth it from ~3 row tiles per CTA on (gate/up: 1792 tiles on
    // 296 CTAs, 21.5 -> 15.8 us at 16 tokens; qkv / o with 384 / 256 tiles lose 0.7-1.4 us to it)
    const int64_t tiles = (g.N + kSkinnyRows - 1) / kSkinnyRows;
    const bool xs = (size_t)skinny_red_bytes<NB>() + (size_t)NB * 8 * (g.K + 16) <= 200 * 1024 && tiles >= 6 * (int64_t)num_sms();
    return xs ? launch_skinny_nb<NB, SILU, true>(g, st) : launch_skinny_nb<NB, SILU, false>(g, st);
}

template <bool SILU>
cudaError_t launch_skinny(const GemvParams& g, cudaStream_t st) {
    if (g.M <= 8) return launch_skinny_xs<1, SILU>(g, st);
    if (g.M <= 16) return launch_skinny_xs<2, SILU>(g, st);
    return launch_skinny_xs<4, SILU>(g, st);
}

template <int MT, int CPL, int R, int MINB, bool SILU = false>
cudaError_t launch_gemv_pre(int64_t ctas, size_t smem, cudaStream_t st, const GemvParams& g) {
    auto kern = gemv_w4a8_pre_kernel<MT, CPL, R, MINB, SILU>;
    static PerDeviceOnce attr;
    if (attr.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int64_t cap = (int64_t)MINB * num_sms();   // one wave of resident CTAs
    return launch_pdl(kern, dim3((unsigned)(cap < ctas ? cap : ctas)), dim3(kGemvWarps * 32), smem, st, g);
}

int launch_rowquant(const void* x, int64_t tokens, int64_t K, float cmax, float cmin, int half_math, uint8_t* xq4,
                    int8_t* xq8, float* sx32, void* sx16, cudaStream_t st) {
    FLATQ_CHECK(tokens > 0 && K > 0 && K % 8 == 0, FLATQ_ERR_INVALID, "quantize_rows: K=%lld must be a multiple of 8",
                (long long)K);
    FLATQ_CHECK(aligned16(x), FLATQ_ERR_INVALID, "x must be 16-byte aligned");
    const int block = 256;
    const int grid = grid_for((size_t)tokens * 32, block, 8);
    FLATQ_CUDA(launch_pdl(rowquant_kernel, dim3(grid), dim3(block), 0, st, reinterpret_cast<const __half*>(x), (int)tokens,
                          (int)K, cmax, cmin, half_math, xq4, xq8, sx32, reinterpret_cast<__half*>(sx16)));
    return FLATQ_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int flatq_abi_version(void) { return FLATQ_ABI_VERSION; }
const char* flatq_last_error(void) { return g_err.c_str(); }

int flatq_device_supported(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

int flatq_pack_i4(const int8_t* q, uint8_t* packed, int64_t rows, int64_t cols, flatq_stream_t stream) {
    FLATQ_CHECK(q && packed, FLATQ_ERR_INVALID, "pack_i4: null pointer");
    FLATQ_CHECK(rows >= 0 && cols >= 0 && cols % 16 == 0, FLATQ_ERR_INVALID, "pack_i4: cols=%lld must be a multiple of 16",
                (long long)cols);
    FLATQ_CHECK(aligned16(q) && (reinterpret_cast<uintptr_t>(packed) & 7) == 0, FLATQ_ERR_INVALID, "pack_i4: alignment");
    const size_t n = (size_t)rows * cols / 16;
    if (n == 0) return FLATQ_OK;
    pack_i4_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(q, packed, n);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_unpack_i4(const uint8_t* packed, int32_t* out, int64_t rows, int64_t cols_packed, flatq_stream_t stream) {
    FLATQ_CHECK(packed && out, FLATQ_ERR_INVALID, "unpack_i4: null pointer");
    FLATQ_CHECK(rows >= 0 && cols_packed >= 0 && cols_packed % 4 == 0, FLATQ_ERR_INVALID,
                "unpack_i4: packed cols=%lld must be a multiple of 4", (long long)cols_packed);
    FLATQ_CHECK((reinterpret_cast<uintptr_t>(packed) & 3) == 0 && aligned16(out), FLATQ_ERR_INVALID, "unpack_i4: alignment");
    const size_t n = (size_t)rows * cols_packed / 4;
    if (n == 0) return FLATQ_OK;
    unpack_i4_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(packed, out, n);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_sym_quant(const void* x, const void* scale, int64_t rows, int64_t cols, uint8_t* q, flatq_stream_t stream) {
    FLATQ_CHECK(x && scale && q, FLATQ_ERR_INVALID, "sym_quant: null pointer");
    FLATQ_CHECK(rows > 0 && cols > 0 && cols % 8 == 0, FLATQ_ERR_INVALID, "sym_quant: cols=%lld must be a multiple of 8",
                (long long)cols);
    FLATQ_CHECK(aligned16(x) && (reinterpret_cast<uintptr_t>(q) & 3) == 0, FLATQ_ERR_INVALID, "sym_quant: alignment");
    const size_t n = (size_t)rows * cols / 8;
    sym_quant_f16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(scale), q, (int)rows, (int)cols);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_sym_dequant(const int32_t* acc, const void* s_row, const void* s_col, const void* bias, int64_t rows,
                      int64_t cols, void* y, int out_dtype, int mode, flatq_stream_t stream) {
    FLATQ_CHECK(acc && s_row && s_col && y, FLATQ_ERR_INVALID, "sym_dequant: null pointer");
    FLATQ_CHECK(rows > 0 && cols > 0 && cols % 4 == 0, FLATQ_ERR_INVALID, "sym_dequant: cols=%lld must be a multiple of 4",
                (long long)cols);
    FLATQ_CHECK(aligned16(acc) && (reinterpret_cast<uintptr_t>(y) & 7) == 0, FLATQ_ERR_INVALID, "sym_dequant: alignment");
    FLATQ_CHECK(mode == 0 || (mode == 1 && out_dtype == FLATQ_F16 && bias == nullptr), FLATQ_ERR_UNSUPPORTED,
                "sym_dequant: mode %d needs fp16 output without bias", mode);
    const size_t n = (size_t)rows * cols / 4;
    const int g = grid_for(n, 256);
    if (out_dtype == FLATQ_F16) {
        sym_dequant_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>(
            acc, reinterpret_cast<const __half*>(s_row), reinterpret_cast<const __half*>(s_col),
            reinterpret_cast<const __half*>(bias), reinterpret_cast<__half*>(y), (int)rows, (int)cols, mode);
    } else if (out_dtype == FLATQ_BF16) {
        sym_dequant_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(
            acc, reinterpret_cast<const __half*>(s_row), reinterpret_cast<const __half*>(s_col),
            reinterpret_cast<const __nv_bfloat16*>(bias), reinterpret_cast<__nv_bfloat16*>(y), (int)rows, (int)cols, mode);
    } else {
        return fail(FLATQ_ERR_INVALID, "sym_dequant: out_dtype %d", out_dtype);
    }
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

size_t flatq_gemm_i4_workspace_bytes(int64_t M, int64_t K) { return align_up((size_t)M * K, 256); }

int flatq_gemm_i4(const uint8_t* A, const uint8_t* B, int64_t M, int64_t N, int64_t K, int32_t* C, void* ws,
                  size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(A && B && C && ws, FLATQ_ERR_INVALID, "gemm_i4: null pointer");
    // the reference asserts packed-K % 32 == 0 (deploy/__init__.py:38)
    FLATQ_CHECK(K % 64 == 0, FLATQ_ERR_INVALID, "gemm_i4: K=%lld must be a multiple of 64", (long long)K);
    FLATQ_CHECK(ws_bytes >= flatq_gemm_i4_workspace_bytes(M, K), FLATQ_ERR_WORKSPACE, "gemm_i4: workspace too small");
    FLATQ_CHECK(aligned16(A) && aligned16(ws), FLATQ_ERR_INVALID, "gemm_i4: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)M * (K / 2) / 16;
    expand_i4_to_i8x16_kernel<<<grid_for(n, 256), 256, 0, st>>>(A, reinterpret_cast<uint8_t*>(ws), n);
    FLATQ_CUDA(cudaGetLastError());
    return launch_gemm(B, ws, M, N, K, 8, C, kEpiI32, nullptr, nullptr, nullptr, st);
}

size_t flatq_linear4bit_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    return align_up((size_t)M * K, 256) + align_up((size_t)M * 4, 256) + align_up((size_t)N * 4, 256);
}

int flatq_linear4bit(const uint8_t* A, const void* s_a, const uint8_t* B, const void* s_b, const void* bias, int64_t M,
                     int64_t N, int64_t K, void* y, int out_dtype, void* ws, size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(A && B && s_a && s_b && y && ws, FLATQ_ERR_INVALID, "linear4bit: null pointer");
    FLATQ_CHECK(K % 64 == 0, FLATQ_ERR_INVALID, "linear4bit: K=%lld must be a multiple of 64", (long long)K);
    FLATQ_CHECK(out_dtype == FLATQ_F16 || out_dtype == FLATQ_BF16, FLATQ_ERR_INVALID, "linear4bit: out_dtype");
    FLATQ_CHECK(ws_bytes >= flatq_linear4bit_workspace_bytes(M, N, K), FLATQ_ERR_WORKSPACE, "linear4bit: workspace too small");
    FLATQ_CHECK(aligned16(A) && aligned16(ws), FLATQ_ERR_INVALID, "linear4bit: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* x8 = reinterpret_cast<uint8_t*>(ws);
    float* sa32 = reinterpret_cast<float*>(x8 + align_up((size_t)M * K, 256));
    float* sb32 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sa32) + align_up((size_t)M * 4, 256));
    const size_t n = (size_t)M * (K / 2) / 16;
    expand_i4_to_i8x16_kernel<<<grid_for(n, 256), 256, 0, st>>>(A, x8, n);
    half_to_float_kernel<<<(int)((M + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(s_a), sa32, (int)M);
    half_to_float_kernel<<<(int)((N + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(s_b), sb32, (int)N);
    FLATQ_CUDA(cudaGetLastError());
    return launch_gemm(B, x8, M, N, K, 8, y, out_dtype == FLATQ_F16 ? kEpiF16 : kEpiBF16, sa32, sb32, bias, st);
}

int flatq_kron_transform(const void* x, const void* left, const void* right, int64_t tokens, int n1, int n2, void* z,
                         flatq_stream_t stream) {
    FLATQ_CHECK(x && z, FLATQ_ERR_INVALID, "kron_transform: null pointer");
    KronParams p = {};
    p.x = reinterpret_cast<const __half*>(x);
    p.left = reinterpret_cast<const __half*>(left);
    p.right = reinterpret_cast<const __half*>(right);
    p.tokens = (int)tokens; p.n1 = n1; p.n2 = n2;
    p.do_quant = 0; p.c_max = p.c_min = 1.f;
    p.z_out = reinterpret_cast<__half*>(z);
    return launch_kron(p, (cudaStream_t)stream);
}

int flatq_transform_quant(const void* x, const void* left, const void* right, int64_t tokens, int n1, int n2,
                          float clip_max, float clip_min, uint8_t* xq_packed, void* scales_f16, float* scales_f32,
                          flatq_stream_t stream) {
    FLATQ_CHECK(x, FLATQ_ERR_INVALID, "transform_quant: null pointer");
    FLATQ_CHECK(!xq_packed || aligned16(xq_packed), FLATQ_ERR_INVALID, "transform_quant: output alignment");
    if (!left && !right)
        return launch_rowquant(x, tokens, (int64_t)n1 * n2, clip_max, clip_min, 0, xq_packed, nullptr, scales_f32,
                               scales_f16, (cudaStream_t)stream);
    KronParams p = {};
    p.x = reinterpret_cast<const __half*>(x);
    p.left = reinterpret_cast<const __half*>(left);
    p.right = reinterpret_cast<const __half*>(right);
    p.tokens = (int)tokens; p.n1 = n1; p.n2 = n2;
    p.do_quant = 1; p.c_max = clip_max; p.c_min = clip_min;
    p.xq4 = xq_packed;
    p.sx_f16 = reinterpret_cast<__half*>(scales_f16);
    p.sx_f32 = scales_f32;
    return launch_kron(p, (cudaStream_t)stream);
}

int flatq_quantize_rows(const void* x, int64_t tokens, int64_t K, float clip_max, float clip_min, int half_math,
                        uint8_t* xq_packed, void* scales_f16, float* scales_f32, flatq_stream_t stream) {
    FLATQ_CHECK(x, FLATQ_ERR_INVALID, "quantize_rows: null pointer");
    return launch_rowquant(x, tokens, K, clip_max, clip_min, half_math, xq_packed, nullptr, scales_f32, scales_f16,
                           (cudaStream_t)stream);
}

// decode shapes (M <= 64): too few 128-row tiles to occupy the chip -> split K over CTAs into an int32 workspace
static int decode_splits(int64_t M, int64_t N, int64_t K) {
    if (M > 64) return 1;
    const int64_t tiles = (N + kGemmBlockW - 1) / kGemmBlockW;
    const int64_t kb = (K + kGemmBlockK - 1) / kGemmBlockK;
    int64_t s = num_sms() / tiles;
    if (s > 8) s = 8;
    if (s > kb) s = kb;
    return s < 2 ? 1 : (int)s;
}

size_t flatq_gemm_w4a8_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    (void)K;
    if (M <= 64) return align_up((size_t)M * N * 4, 256);          // split-K accumulator (decode)
    return streamk_ws_bytes(M <= 128 ? 128 : 256);                  // stream-K partials + flags (prefill)
}

size_t flatq_linear_fused_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    return align_up((size_t)M * K, 256) + align_up((size_t)M * 4, 256) + flatq_gemm_w4a8_workspace_bytes(M, N, K);
}

int flatq_act_quant(const void* x, const void* left, const void* right, int n1, int n2, float clip_max, float clip_min,
                    int64_t M, int8_t* xq8, float* s_x, flatq_stream_t stream) {
    FLATQ_CHECK(x && xq8 && s_x, FLATQ_ERR_INVALID, "act_quant: null pointer");
    FLATQ_CHECK(aligned16(xq8), FLATQ_ERR_INVALID, "act_quant: xq8 alignment");
    const int64_t K = (int64_t)n1 * n2;
    FLATQ_CHECK(K % 32 == 0, FLATQ_ERR_INVALID, "act_quant: K=%lld must be a multiple of 32", (long long)K);
    cudaStream_t st = (cudaStream_t)stream;
    if (!left && !right) return launch_rowquant(x, M, K, clip_max, clip_min, 0, nullptr, xq8, s_x, nullptr, st);
    KronParams p = {};
    p.x = reinterpret_cast<const __half*>(x);
    p.left = reinterpret_cast<const __half*>(left);
    p.right = reinterpret_cast<const __half*>(right);
    p.tokens = (int)M; p.n1 = n1; p.n2 = n2;
    p.do_quant = 1; p.c_max = clip_max; p.c_min = clip_min;
    p.xq8 = xq8; p.sx_f32 = s_x;
    return launch_kron(p, st);
}

int flatq_gemm_w4a8(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, const void* bias,
                    int64_t M, int64_t N, int64_t K, void* y, int out_dtype, void* ws, size_t ws_bytes,
                    flatq_stream_t stream) {
    FLATQ_CHECK(xq8 && s_x && w_packed && s_w && y, FLATQ_ERR_INVALID, "gemm_w4a8: null pointer");
    FLATQ_CHECK(out_dtype == FLATQ_F16 || out_dtype == FLATQ_BF16, FLATQ_ERR_INVALID, "gemm_w4a8: out_dtype");
    FLATQ_CHECK(K % 32 == 0, FLATQ_ERR_INVALID, "gemm_w4a8: K=%lld must be a multiple of 32", (long long)K);
    cudaStream_t st = (cudaStream_t)stream;
    if (M <= 2 && (size_t)M * K <= 96 * 1024) {  // decode: dp4a GEMV streams the packed weight at HBM speed
        FLATQ_CHECK(aligned16(xq8) && aligned16(w_packed), FLATQ_ERR_INVALID, "gemm_w4a8: alignment");
        GemvParams g;
        g.w = w_packed; g.x = xq8; g.s_x = s_x; g.s_w = s_w; g.bias = bias; g.out = y;
        g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bf16_out = out_dtype == FLATQ_BF16;
        int64_t ctas = (N + kGemvWarps - 1) / kGemvWarps;
        const int64_t cap = (int64_t)num_sms() * 6;
        const dim3 grid((unsigned)(ctas < cap ? ctas : cap)), block(kGemvWarps * 32);
        const size_t smem = (size_t)(M == 1 ? 1 : 2) * K;
        auto k1 = gemv_w4a8_kernel<1>;
        auto k2 = gemv_w4a8_kernel<2>;
        static PerDeviceOnce attr;
        if (attr.first_time()) {
            FLATQ_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            FLATQ_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
        static const bool generic = getenv("FLATQ_GEMV_GENERIC") != nullptr;   // A/B aid
        if (K % 1024 == 0 && !generic) {   // weight rows in registers, first rows requested before the PDL wait
            switch ((int)(K / 1024) * 10 + (int)M) {
#define FLATQ_GEMV_PRE(CPL, R, MINB) /* R rows of CPL 16-byte chunks per lane in registers, MINB CTAs per SM */           \
    case CPL * 10 + 1: FLATQ_CUDA((launch_gemv_pre<1, CPL, R, MINB>(ctas, smem, st, g))); return FLATQ_OK;              \
    case CPL * 10 + 2: FLATQ_CUDA((launch_gemv_pre<2, CPL, R, MINB>(ctas, smem, st, g))); return FLATQ_OK;
                FLATQ_GEMV_PRE(4, 1, 5)     // K = 4096: short rows, many CTAs (as the generic kernel), first row before the wait
                FLATQ_GEMV_PRE(5, 1, 4)     // 5120
                FLATQ_GEMV_PRE(7, 1, 4)     // 7168
                FLATQ_GEMV_PRE(8, 1, 4)     // 8192
                FLATQ_GEMV_PRE(14, 1, 2)    // 14336: a row is 56 registers per lane
#undef FLATQ_GEMV_PRE
                default: break;
            }
        }
        FLATQ_CUDA(M == 1 ? launch_pdl(k1, grid, block, smem, st, g) : launch_pdl(k2, grid, block, smem, st, g));
        return FLATQ_OK;
    }
    if (skinny_ok(M, N, K, xq8, w_packed)) {
        GemvParams g;
        g.w = w_packed; g.x = xq8; g.s_x = s_x; g.s_w = s_w; g.bias = bias; g.out = y;
        g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bf16_out = out_dtype == FLATQ_BF16;
        FLATQ_CUDA(launch_skinny<false>(g, st));
        return FLATQ_OK;
    }
    const int splits = (ws && ws_bytes >= flatq_gemm_w4a8_workspace_bytes(M, N, K)) ? decode_splits(M, N, K) : 1;
    if (splits == 1)
        return launch_gemm(w_packed, xq8, M, N, K, 4, y, out_dtype == FLATQ_F16 ? kEpiF16 : kEpiBF16, s_x, s_w, bias, st, 1,
                           M > 64 ? ws : nullptr, M > 64 ? ws_bytes : 0);
    // split-K: zero the int32 workspace, accumulate partial tiles with atomics, dequantise in a finalize pass
    FLATQ_CHECK(aligned16(ws), FLATQ_ERR_INVALID, "gemm_w4a8: workspace alignment");
    FLATQ_CUDA(cudaMemsetAsync(ws, 0, (size_t)M * N * 4, st));
    int rc = launch_gemm(w_packed, xq8, M, N, K, 4, ws, kEpiI32, nullptr, nullptr, nullptr, st, splits);
    if (rc) return rc;
    const int g = grid_for((size_t)M * N, 256, 4);
    if (out_dtype == FLATQ_F16)
        splitk_finalize_kernel<__half><<<g, 256, 0, st>>>(reinterpret_cast<const int32_t*>(ws), s_x, s_w,
                                                          reinterpret_cast<const __half*>(bias), reinterpret_cast<__half*>(y), (int)M, (int)N);
    else
        splitk_finalize_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(reinterpret_cast<const int32_t*>(ws), s_x, s_w,
                                                                 reinterpret_cast<const __nv_bfloat16*>(bias),
                                                                 reinterpret_cast<__nv_bfloat16*>(y), (int)M, (int)N);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

// ---- fused [gate; up] GEMM + SiLU(gate) * up (row-interleaved weight)
static bool silu_in_epilogue(int64_t M) { return M > 128 && !opt_gemm_1sm(); }

size_t flatq_gemm_w4a8_silu_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (silu_in_epilogue(M)) return 0;
    return align_up((size_t)M * N * 2, 256) + flatq_gemm_w4a8_workspace_bytes(M, N, K);
}

int flatq_gemm_w4a8_silu(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, int64_t M,
                         int64_t N, int64_t K, void* out_f16, void* ws, size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(xq8 && s_x && w_packed && s_w && out_f16, FLATQ_ERR_INVALID, "gemm_w4a8_silu: null pointer");
    FLATQ_CHECK(N > 0 && N % 16 == 0, FLATQ_ERR_INVALID, "gemm_w4a8_silu: N=%lld (2 x intermediate size) must be a multiple of 16", (long long)N);
    FLATQ_CHECK(K % 32 == 0, FLATQ_ERR_INVALID, "gemm_w4a8_silu: K=%lld must be a multiple of 32", (long long)K);
    cudaStream_t st = (cudaStream_t)stream;
    if (silu_in_epilogue(M))
        return launch_gemm(w_packed, xq8, M, N, K, 4, out_f16, kEpiF16, s_x, s_w, nullptr, st, 1, nullptr, 0, true);
    if (M <= 2 && (size_t)M * K <= 96 * 1024 && K % 1024 == 0 && aligned16(xq8) && aligned16(w_packed)) {
        // decode: the GEMV's warps own (gate, up) row pairs and apply the activation themselves -- one launch
        GemvParams g;
        g.w = w_packed; g.x = xq8; g.s_x = s_x; g.s_w = s_w; g.bias = nullptr; g.out = out_f16;
        g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bf16_out = 0;
        const int64_t ctas = (N / 2 + kGemvWarps - 1) / kGemvWarps;
        const size_t smem = (size_t)(M == 1 ? 1 : 2) * K;
        switch ((int)(K / 1024) * 10 + (int)M) {
#define FLATQ_GEMV_SILU(CPL, MINB)                                                                                      \
    case CPL * 10 + 1: FLATQ_CUDA((launch_gemv_pre<1, CPL, 2, MINB, true>(ctas, smem, st, g))); return FLATQ_OK;        \
    case CPL * 10 + 2: FLATQ_CUDA((launch_gemv_pre<2, CPL, 2, MINB, true>(ctas, smem, st, g))); return FLATQ_OK;
            FLATQ_GEMV_SILU(4, 3)    // hidden 4096
            FLATQ_GEMV_SILU(5, 2)    // 5120
            FLATQ_GEMV_SILU(8, 1)    // 8192
#undef FLATQ_GEMV_SILU
            default: break;
        }
    }
    if (skinny_ok(M, N, K, xq8, w_packed)) {   // decode batches: the activation in the streaming kernel's epilogue
        GemvParams g;
        g.w = w_packed; g.x = xq8; g.s_x = s_x; g.s_w = s_w; g.bias = nullptr; g.out = out_f16;
        g.M = (int)M; g.N = (int)N; g.K = (int)K; g.bf16_out = 0;
        FLATQ_CUDA(launch_skinny<true>(g, st));
        return FLATQ_OK;
    }
    FLATQ_CHECK(ws && aligned16(ws) && ws_bytes >= flatq_gemm_w4a8_silu_workspace_bytes(M, N, K), FLATQ_ERR_WORKSPACE,
                "gemm_w4a8_silu: workspace too small or misaligned");
    FLATQ_CHECK(aligned16(out_f16), FLATQ_ERR_INVALID, "gemm_w4a8_silu: output alignment");
    const size_t off = align_up((size_t)M * N * 2, 256);
    int rc = flatq_gemm_w4a8(xq8, s_x, w_packed, s_w, nullptr, M, N, K, ws, FLATQ_F16, reinterpret_cast<uint8_t*>(ws) + off,
                             ws_bytes - off, stream);
    if (rc) return rc;
    const size_t total_vec = (size_t)M * N / 16;
    FLATQ_CUDA(launch_pdl(silu_mul_il_f16_kernel, dim3(grid_for(total_vec, 256, 4)), dim3(256), 0, st,
                          reinterpret_cast<const __half*>(ws), reinterpret_cast<__half*>(out_f16), total_vec));
    return FLATQ_OK;
}

size_t flatq_gemm_w4a8_grouped_workspace_bytes(int n_groups) {
    return align_up((size_t)(n_groups > 0 ? n_groups + 1 : 1) * sizeof(int32_t), 256);
}

int flatq_gemm_w4a8_grouped(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w,
                            const void* bias, const int32_t* group_offsets, int n_groups, int64_t M, int64_t N, int64_t K,
                            int rows_per_group_hint, void* y, int out_dtype, void* ws, size_t ws_bytes,
                            flatq_stream_t stream) {
    FLATQ_CHECK(xq8 && s_x && w_packed && s_w && group_offsets && y && ws, FLATQ_ERR_INVALID, "gemm_w4a8_grouped: null pointer");
    FLATQ_CHECK(out_dtype == FLATQ_F16 || out_dtype == FLATQ_BF16, FLATQ_ERR_INVALID, "gemm_w4a8_grouped: out_dtype");
    FLATQ_CHECK(n_groups >= 1 && n_groups <= 65536, FLATQ_ERR_INVALID, "gemm_w4a8_grouped: n_groups=%d", n_groups);
    FLATQ_CHECK(M > 0 && N > 0 && K > 0, FLATQ_ERR_INVALID, "gemm_w4a8_grouped: empty GEMM %lldx%lldx%lld", (long long)M,
                (long long)N, (long long)K);
    FLATQ_CHECK(K % 32 == 0 && K <= 65536, FLATQ_ERR_INVALID, "gemm_w4a8_grouped: K=%lld must be a multiple of 32, <= 65536",
                (long long)K);
    FLATQ_CHECK(N % 2 == 0, FLATQ_ERR_UNSUPPORTED, "gemm_w4a8_grouped: N=%lld must be even", (long long)N);
    FLATQ_CHECK((M + 256) * N < (1ll << 32) && (int64_t)n_groups * N < (1ll << 31), FLATQ_ERR_UNSUPPORTED,
                "gemm_w4a8_grouped: output / stacked weight exceeds the 32-bit index of the store path");
    FLATQ_CHECK(ws_bytes >= flatq_gemm_w4a8_grouped_workspace_bytes(n_groups) && (reinterpret_cast<uintptr_t>(ws) & 3) == 0,
                FLATQ_ERR_WORKSPACE, "gemm_w4a8_grouped: workspace too small or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t hint = rows_per_group_hint > 0 ? rows_per_group_hint : (M + n_groups - 1) / n_groups;
    // Token tile ~ twice the mean group size: below N = 128 a tcgen05.mma sits on its ~45-cycle issue floor (DESIGN.md
    // section 2), so a wider tile costs little tensor time and halves the MMA count of every group above the mean
    // (routing is skewed); 256 (single accumulator buffer, 3 stages) only when the mean itself exceeds 128.
    const int tn = hint <= 8 ? 16 : hint <= 16 ? 32 : hint <= 32 ? 64 : hint <= 128 ? 128 : 256;
    const int64_t nb_w = (N + kGemmBlockW - 1) / kGemmBlockW;
    const int64_t t_upper = nb_w * ((M + tn - 1) / tn + n_groups);  // every group adds at most one partial token block
    FLATQ_CHECK(t_upper < (1ll << 31), FLATQ_ERR_UNSUPPORTED, "gemm_w4a8_grouped: too many tiles");
    CUtensorMap tw, tx;
    int rc = make_map_u8(&tw, w_packed, (uint64_t)n_groups * N, K / 2, K / 2, kGemmBlockK / 2, kGemmBlockW, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_u8(&tx, xq8, M, K, K, 128, tn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    int* tile_prefix = reinterpret_cast<int*>(ws);
    grouped_prefix_kernel<<<1, kPrefixThreads, 0, st>>>(group_offsets, n_groups, tn, (int)nb_w, tile_prefix);
    FLATQ_CUDA(cudaGetLastError());
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.acc_shift = 4;
    p.out = y; p.s_x = s_x; p.s_w = s_w; p.bias = bias; p.prof = nullptr; p.dbg = 0; p.k_splits = 1;
    p.streamk = 0; p.sk_part = nullptr; p.sk_flags = nullptr; p.epoch = 0;
    p.grp_off = group_offsets; p.grp_tile = tile_prefix; p.n_groups = n_groups; p.t_upper = (int)t_upper;
    p.rows_per_owner = 0; p.tp_rank = 0; p.silu_pairs = 0;
    return out_dtype == FLATQ_F16 ? launch_gemm_grouped_m<kEpiF16>(tn, tw, tx, p, st)
                                  : launch_gemm_grouped_m<kEpiBF16>(tn, tw, tx, p, st);
}

// ---- tensor-parallel row-parallel linears over NVLink peer memory (csrc/peer_reduce.cuh)
static int peer_ctx(const flatq_tp_ctx* tp, PeerCtx* c) {
    FLATQ_CHECK(tp && tp->world >= 2 && tp->world <= kPeerMaxWorld && tp->rank >= 0 && tp->rank < tp->world, FLATQ_ERR_INVALID,
                "tp: world must be 2..%d and 0 <= rank < world", kPeerMaxWorld);
    c->rank = tp->rank; c->world = tp->world;
    c->pad_ctas = (int)(tp->signal_pad_bytes / (2ull * (tp->world + 1) * sizeof(uint32_t)));   // two phases x (world flags + 1 counter) per CTA slot
    FLATQ_CHECK(c->pad_ctas >= 1, FLATQ_ERR_INVALID, "tp: signal pad too small (%zu bytes)", tp->signal_pad_bytes);
    for (int r = 0; r < tp->world; ++r) {
        FLATQ_CHECK(tp->signal_pads[r], FLATQ_ERR_INVALID, "tp: null signal pad for rank %d", r);
        c->signal_pads[r] = reinterpret_cast<uint32_t*>(tp->signal_pads[r]);
    }
    return FLATQ_OK;
}

static int peer_grid(const flatq_tp_ctx* tp, long long vecs) {
    long long g = (vecs + kPeerThreads * 4 - 1) / (kPeerThreads * 4);
    const long long cap = tp->signal_pad_bytes / (2ull * (tp->world + 1) * sizeof(uint32_t));   // two sets of grid x (world flags + 1 counter)
    const long long lim = cap < 96 ? cap : 96;   // measured at tp 2: 96 CTAs 46.5 us, 296 CTAs 56.6 us for 16 MiB (every CTA runs its own cross-GPU barrier)
    if (g > lim) g = lim;
    return (int)(g < 1 ? 1 : g);
}

int flatq_gemm_w4a8_rowparallel(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, int64_t M,
                                int64_t N, int64_t K, const flatq_tp_ctx* tp, void* const* peer_staging, flatq_stream_t stream) {
    FLATQ_CHECK(xq8 && s_x && w_packed && s_w && tp && peer_staging, FLATQ_ERR_INVALID, "gemm_w4a8_rowparallel: null pointer");
    FLATQ_CHECK(tp->world >= 2 && tp->world <= kPeerMaxWorld, FLATQ_ERR_INVALID, "gemm_w4a8_rowparallel: world");
    FLATQ_CHECK(M % (256ll * tp->world) == 0, FLATQ_ERR_UNSUPPORTED,
                "gemm_w4a8_rowparallel: M=%lld must be a multiple of 256 x world (use flatq_gemm_w4a8 + flatq_peer_allreduce)", (long long)M);
    FLATQ_CHECK(K % 32 == 0 && N % 8 == 0 && (M + 256) * N < (1ll << 32), FLATQ_ERR_INVALID, "gemm_w4a8_rowparallel: shape");
    CUtensorMap tw, tx;
    int rc = make_map_u8(&tw, w_packed, N, K / 2, K / 2, kGemmBlockK / 2, kGemmBlockW, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_u8(&tx, xq8, M, K, K, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.acc_shift = 4;
    p.out = nullptr; p.s_x = s_x; p.s_w = s_w; p.bias = nullptr; p.prof = nullptr; p.dbg = 0; p.k_splits = 1;
    p.streamk = 0; p.sk_part = nullptr; p.sk_flags = nullptr; p.epoch = 0;
    p.grp_off = nullptr; p.grp_tile = nullptr; p.n_groups = 0; p.t_upper = 0;
    p.rows_per_owner = (int)(M / tp->world); p.tp_rank = tp->rank; p.silu_pairs = 0;
    for (int r = 0; r < 8; ++r) p.peer_out[r] = r < tp->world ? peer_staging[r] : nullptr;
    for (int r = 0; r < tp->world; ++r) FLATQ_CHECK(p.peer_out[r], FLATQ_ERR_INVALID, "gemm_w4a8_rowparallel: null staging pointer");
    return launch_gemm_2sm<kEpiF16>(tw, tx, p, (cudaStream_t)stream);
}

int flatq_peer_reduce_bcast(const flatq_tp_ctx* tp, const void* staging, void* out_mc, void* const* out_peers, const void* bias,
                            int64_t M, int64_t N, flatq_stream_t stream) {
    PeerCtx c;
    int rc = peer_ctx(tp, &c);
    if (rc) return rc;
    FLATQ_CHECK(staging && (out_mc || out_peers), FLATQ_ERR_INVALID, "peer_reduce_bcast: null pointer");
    FLATQ_CHECK(M % tp->world == 0 && N % 8 == 0 && aligned16(staging) && (!bias || aligned16(bias)), FLATQ_ERR_INVALID, "peer_reduce_bcast: shape / alignment");
    PeerBcastDst dst;
    for (int r = 0; r < kPeerMaxWorld; ++r) dst.peers[r] = (out_peers && r < tp->world) ? reinterpret_cast<__half*>(out_peers[r]) : nullptr;
    const int rpo = (int)(M / tp->world);
    const int grid = peer_grid(tp, (long long)rpo * N / 8);
    peer_reduce_bcast_kernel<<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(c, reinterpret_cast<const __half*>(staging),
                                                                             reinterpret_cast<__half*>(out_mc), dst,
                                                                             reinterpret_cast<const __half*>(bias), rpo, (int)N);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_peer_allreduce(const flatq_tp_ctx* tp, const void* src_mc, void* dst_mc, const void* bias, int64_t rows, int64_t N,
                         flatq_stream_t stream) {
    PeerCtx c;
    int rc = peer_ctx(tp, &c);
    if (rc) return rc;
    FLATQ_CHECK(src_mc && dst_mc, FLATQ_ERR_INVALID, "peer_allreduce: needs the multicast mapping of both buffers");
    FLATQ_CHECK(rows > 0 && N % 8 == 0 && aligned16(src_mc) && aligned16(dst_mc), FLATQ_ERR_INVALID, "peer_allreduce: shape / alignment");
    const int grid = peer_grid(tp, rows * N / 8 / tp->world);
    peer_allreduce_nvls_kernel<<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(c, reinterpret_cast<const __half*>(src_mc),
                                                                               reinterpret_cast<__half*>(dst_mc),
                                                                               reinterpret_cast<const __half*>(bias), rows, (int)N);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_linear_fused(const void* x, const void* left, const void* right, int n1, int n2, float clip_max,
                       float clip_min, const uint8_t* w_packed, const float* s_w, const void* bias, int64_t M, int64_t N,
                       void* y, int out_dtype, void* ws, size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(x && w_packed && s_w && y && ws, FLATQ_ERR_INVALID, "linear_fused: null pointer");
    const int64_t K = (int64_t)n1 * n2;
    if (N > 0 && N < (1ll << 31) && flat_decode_ok(M, N, n1, n2, x, left, right, w_packed, false)) {   // one launch for 1-2 tokens
        FLATQ_CHECK(out_dtype == FLATQ_F16 || out_dtype == FLATQ_BF16, FLATQ_ERR_INVALID, "linear_fused: out_dtype");
        FlatDecodeParams q;
        q.g.w = w_packed; q.g.x = nullptr; q.g.s_x = nullptr; q.g.s_w = s_w; q.g.bias = bias; q.g.out = y;
        q.g.M = (int)M; q.g.N = (int)N; q.g.K = (int)K; q.g.bf16_out = out_dtype == FLATQ_BF16;
        q.x = reinterpret_cast<const __half*>(x); q.left = reinterpret_cast<const __half*>(left);
        q.right = reinterpret_cast<const __half*>(right); q.n1 = n1; q.n2 = n2; q.c_max = clip_max; q.c_min = clip_min;
        q.dbg_codes = g_decode_dump_codes; q.dbg_sx = g_decode_dump_sx;
        FLATQ_CUDA(launch_flat_decode<false>(q, (cudaStream_t)stream));
        return FLATQ_OK;
    }
    FLATQ_CHECK(ws_bytes >= flatq_linear_fused_workspace_bytes(M, N, K), FLATQ_ERR_WORKSPACE, "linear_fused: workspace too small");
    FLATQ_CHECK(aligned16(ws), FLATQ_ERR_INVALID, "linear_fused: workspace alignment");
    int8_t* x8 = reinterpret_cast<int8_t*>(ws);
    uint8_t* p = reinterpret_cast<uint8_t*>(ws) + align_up((size_t)M * K, 256);
    float* sx = reinterpret_cast<float*>(p);
    void* acc_ws = p + align_up((size_t)M * 4, 256);  // split-K accumulator (decode) or stream-K partials (prefill)
    int rc = flatq_act_quant(x, left, right, n1, n2, clip_max, clip_min, M, x8, sx, stream);
    if (rc) return rc;
    return flatq_gemm_w4a8(x8, sx, w_packed, s_w, bias, M, N, K, y, out_dtype, acc_ws,
                           flatq_gemm_w4a8_workspace_bytes(M, N, K), stream);
}

size_t flatq_linear_fused_silu_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    return align_up((size_t)M * K, 256) + align_up((size_t)M * 4, 256) + flatq_gemm_w4a8_silu_workspace_bytes(M, N, K);
}

int flatq_linear_fused_silu(const void* x, const void* left, const void* right, int n1, int n2, float clip_max, float clip_min,
                            const uint8_t* w_packed_il, const float* s_w_il, int64_t M, int64_t N, void* out_f16, void* ws,
                            size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(x && w_packed_il && s_w_il && out_f16, FLATQ_ERR_INVALID, "linear_fused_silu: null pointer");
    FLATQ_CHECK(N > 0 && N % 16 == 0 && N < (1ll << 31), FLATQ_ERR_INVALID, "linear_fused_silu: N=%lld (2 x intermediate size) must be a multiple of 16", (long long)N);
    const int64_t K = (int64_t)n1 * n2;
    if (flat_decode_ok(M, N, n1, n2, x, left, right, w_packed_il, true)) {
        FlatDecodeParams q;
        q.g.w = w_packed_il; q.g.x = nullptr; q.g.s_x = nullptr; q.g.s_w = s_w_il; q.g.bias = nullptr; q.g.out = out_f16;
        q.g.M = (int)M; q.g.N = (int)N; q.g.K = (int)K; q.g.bf16_out = 0;
        q.x = reinterpret_cast<const __half*>(x); q.left = reinterpret_cast<const __half*>(left);
        q.right = reinterpret_cast<const __half*>(right); q.n1 = n1; q.n2 = n2; q.c_max = clip_max; q.c_min = clip_min;
        q.dbg_codes = g_decode_dump_codes; q.dbg_sx = g_decode_dump_sx;
        FLATQ_CUDA(launch_flat_decode<true>(q, (cudaStream_t)stream));
        return FLATQ_OK;
    }
    FLATQ_CHECK(ws && aligned16(ws) && ws_bytes >= flatq_linear_fused_silu_workspace_bytes(M, N, K), FLATQ_ERR_WORKSPACE,
                "linear_fused_silu: workspace too small or misaligned");
    int8_t* x8 = reinterpret_cast<int8_t*>(ws);
    uint8_t* p = reinterpret_cast<uint8_t*>(ws) + align_up((size_t)M * K, 256);
    float* sx = reinterpret_cast<float*>(p);
    uint8_t* rest = p + align_up((size_t)M * 4, 256);
    int rc = flatq_act_quant(x, left, right, n1, n2, clip_max, clip_min, M, x8, sx, stream);
    if (rc) return rc;
    return flatq_gemm_w4a8_silu(x8, sx, w_packed_il, s_w_il, M, N, K, out_f16, rest,
                                ws_bytes - (size_t)(rest - reinterpret_cast<uint8_t*>(ws)), stream);
}

int flatq_hadamard(const void* x, void* out, int64_t rows, int n, float scale, flatq_stream_t stream) {
    FLATQ_CHECK(x && out, FLATQ_ERR_INVALID, "hadamard: null pointer");
    FLATQ_CHECK(n >= 8 && n <= 32768 && (n & (n - 1)) == 0, FLATQ_ERR_UNSUPPORTED,
                "hadamard: n=%d must be a power of two in [8, 32768]", n);
    FLATQ_CHECK(rows > 0 && aligned16(x) && aligned16(out), FLATQ_ERR_INVALID, "hadamard: rows / alignment");
    int log_n = 0;
    while ((1 << log_n) < n) ++log_n;
    if (n <= 256) {   // a row in 1..32 lanes of a warp, no shared memory
        const int lanes = n / 8;
        const int64_t warps = (rows + 32 / lanes - 1) / (32 / lanes);
        const dim3 grid((unsigned)grid_for((size_t)warps * 32, 256, 8)), block(256);
        const __half* xi = reinterpret_cast<const __half*>(x);
        __half* oo = reinterpret_cast<__half*>(out);
        cudaStream_t st = (cudaStream_t)stream;
        switch (lanes) {
            case 1: fht_warp_kernel<1><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
            case 2: fht_warp_kernel<2><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
            case 4: fht_warp_kernel<4><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
            case 8: fht_warp_kernel<8><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
            case 16: fht_warp_kernel<16><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
            default: fht_warp_kernel<32><<<grid, block, 0, st>>>(xi, oo, rows, scale); break;
        }
        FLATQ_CUDA(cudaGetLastError());
        return FLATQ_OK;
    }
    // threads per row: one per 8 elements (at least 8); short rows share a CTA of 256 threads
    const int tpr = n / 8 < 8 ? 8 : (n / 8 > 1024 ? 1024 : n / 8);
    int rpc = tpr >= 256 ? 1 : 256 / tpr;
    if (rows < rpc) rpc = (int)rows;
    const int threads = tpr * rpc;
    const size_t smem = (size_t)rpc * (n + n / 32 + 1) * sizeof(float);
    static PerDeviceMax attr;
    if (smem > 48 * 1024 && attr.grows(smem))
        FLATQ_CUDA(cudaFuncSetAttribute(fht_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ctas = (rows + rpc - 1) / rpc;
    // one wave of resident CTAs: several rows per SM in flight hide each row's load -> passes -> store chain
    int64_t per_sm = 2048 / threads;
    if ((int64_t)(200 * 1024 / (smem + 1024)) < per_sm) per_sm = (int64_t)(200 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    fht_kernel<<<(int)(ctas < cap ? ctas : cap), threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), (int)rows, n, log_n, scale, tpr);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_hadamard_kn(const void* x, void* out, const void* hadK, int64_t rows, int n, int K, float scale,
                      flatq_stream_t stream) {
    FLATQ_CHECK(x && out && hadK, FLATQ_ERR_INVALID, "hadamard_kn: null pointer");
    FLATQ_CHECK(K >= 4 && K <= 172 && K % 4 == 0 && n % K == 0, FLATQ_ERR_UNSUPPORTED,
                "hadamard_kn: K=%d must be a multiple of 4 in [4, 172] dividing n=%d", K, n);
    const int n2 = n / K;
    FLATQ_CHECK(n2 >= 8 && (n2 & (n2 - 1)) == 0 && n <= 32768, FLATQ_ERR_UNSUPPORTED,
                "hadamard_kn: n/K=%d must be a power of two >= 8, n <= 32768", n2);
    FLATQ_CHECK(rows > 0 && aligned16(x) && aligned16(out), FLATQ_ERR_INVALID, "hadamard_kn: rows / alignment");
    int log_n2 = 0;
    while ((1 << log_n2) < n2) ++log_n2;
    const int threads = n / 8 < 128 ? 128 : (n / 8 > 1024 ? 1024 : ((n / 8 + 31) / 32) * 32);
    const size_t smem = (size_t)(n + n / 32 + 2) * sizeof(float) + (size_t)K * K * 2;
    FLATQ_CHECK(smem <= 227 * 1024, FLATQ_ERR_UNSUPPORTED, "hadamard_kn: row of n=%d with K=%d does not fit shared memory", n, K);
    static PerDeviceMax attr;
    if (smem > 48 * 1024 && attr.grows(smem))
        FLATQ_CUDA(cudaFuncSetAttribute(fht_kn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)num_sms() * (smem <= 56 * 1024 ? 4 : (smem <= 113 * 1024 ? 2 : 1));
    fht_kn_kernel<<<(int)(rows < cap ? rows : cap), threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), reinterpret_cast<const __half*>(hadK), (int)rows,
        K, log_n2, scale);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_kv_quant(const void* x, int64_t tokens, int heads, int head_dim, uint8_t* q_out, void* param_out, uint8_t* kv_data,
                   void* kv_param, const int32_t* slot, int layer, int n_layers, int kv_select, int page_size,
                   flatq_stream_t stream) {
    FLATQ_CHECK(x && (q_out || kv_data), FLATQ_ERR_INVALID, "kv_quant: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "kv_quant: head_dim=%d (only 128, as in the reference bindings.cpp:148)", head_dim);
    FLATQ_CHECK(!kv_data || (kv_param && slot && page_size > 0 && n_layers > 0 && layer >= 0 && layer < n_layers &&
                             (kv_select == 0 || kv_select == 1)), FLATQ_ERR_INVALID, "kv_quant: paged-cache arguments");
    FLATQ_CHECK(tokens > 0 && heads > 0, FLATQ_ERR_INVALID, "kv_quant: empty");
    KvQuantParams p;
    p.x = reinterpret_cast<const __half*>(x); p.tokens = (int)tokens; p.heads = heads; p.D = head_dim;
    p.q_out = q_out; p.param_out = reinterpret_cast<__half*>(param_out);
    p.kv_data = kv_data; p.kv_param = reinterpret_cast<__half*>(kv_param); p.slot = slot;
    p.layer = layer; p.n_layers = n_layers; p.kv_select = kv_select; p.page_size = page_size;
    kv_quant_kernel<<<grid_for((size_t)tokens * heads * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(p);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

int flatq_kv_fused(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t tokens,
                   int q_heads, int kv_heads, int head_dim, const void* cos_t, const void* sin_t, const int32_t* pos,
                   const void* t_q, const void* t_k, void* q_out, void* k_out, uint8_t* kv_data, void* kv_param,
                   const int32_t* slot, int layer, int n_layers, int page_size, flatq_stream_t stream) {
    FLATQ_CHECK(q && k && v && q_out && kv_data && kv_param && slot, FLATQ_ERR_INVALID, "kv_fused: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "kv_fused: head_dim=%d (only 128, as in the reference bindings.cpp:148)", head_dim);
    FLATQ_CHECK(tokens > 0 && q_heads > 0 && kv_heads > 0 && tokens * q_heads < (1ll << 31), FLATQ_ERR_INVALID, "kv_fused: tokens / heads");
    FLATQ_CHECK((cos_t == nullptr) == (sin_t == nullptr) && (!cos_t || pos), FLATQ_ERR_INVALID, "kv_fused: RoPE needs cos, sin and pos");
    FLATQ_CHECK(page_size > 0 && n_layers > 0 && layer >= 0 && layer < n_layers, FLATQ_ERR_INVALID, "kv_fused: paged-cache arguments");
    FLATQ_CHECK(ld_q % 4 == 0 && ld_k % 4 == 0 && ld_v % 4 == 0 && ld_q >= (int64_t)q_heads * 128 && ld_k >= (int64_t)kv_heads * 128 &&
                ld_v >= (int64_t)kv_heads * 128, FLATQ_ERR_INVALID, "kv_fused: row strides");
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    FLATQ_CHECK(al8(q) && al8(k) && al8(v) && aligned16(q_out) && (!k_out || aligned16(k_out)) && (!cos_t || (al8(cos_t) && al8(sin_t))) &&
                (!t_q || aligned16(t_q)) && (!t_k || aligned16(t_k)), FLATQ_ERR_INVALID, "kv_fused: alignment");
    KvFusedParams p;
    p.q = reinterpret_cast<const __half*>(q); p.k = reinterpret_cast<const __half*>(k); p.v = reinterpret_cast<const __half*>(v);
    p.ld_q = ld_q; p.ld_k = ld_k; p.ld_v = ld_v;
    p.tokens = (int)tokens; p.q_heads = q_heads; p.kv_heads = kv_heads;
    p.cos_t = reinterpret_cast<const __half*>(cos_t); p.sin_t = reinterpret_cast<const __half*>(sin_t); p.pos = pos;
    p.t_q = reinterpret_cast<const __half*>(t_q); p.t_k = reinterpret_cast<const __half*>(t_k);
    p.q_out = reinterpret_cast<__half*>(q_out); p.k_out = reinterpret_cast<__half*>(k_out);
    p.kv_data = kv_data; p.kv_param = reinterpret_cast<__half*>(kv_param); p.slot = slot;
    p.layer = layer; p.n_layers = n_layers; p.page_size = page_size;
    // a decode step has a few hundred rows: spread them over the chip in 2-warp CTAs instead of stacking 8 latency chains
    // on each of a handful of SMs (16 tokens of Llama-3-8B: 6 CTAs of 8 warps -> 24 of 2; measured 19 -> see DESIGN.md 6)
    const int64_t tiles = (tokens * (q_heads + 2 * (int64_t)kv_heads) + 15) / 16;
    const int warps = tiles <= 2 * num_sms() ? 2 : tiles <= 4 * num_sms() ? 4 : 8;
    p.rows_per_cta = 16 * warps;
    p.q_blocks = (int)((tokens * q_heads + p.rows_per_cta - 1) / p.rows_per_cta);
    p.k_blocks = (int)((tokens * kv_heads + p.rows_per_cta - 1) / p.rows_per_cta);
    static PerDeviceOnce attr;
    if (attr.first_time()) FLATQ_CUDA(cudaFuncSetAttribute(kv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvfSmem));
    FLATQ_CUDA(launch_pdl(kv_fused_kernel, dim3(p.q_blocks + 2 * p.k_blocks), dim3(32 * warps), (size_t)kKvfSmem, (cudaStream_t)stream, p));
    return FLATQ_OK;
}

int flatq_kv_decode_i4(const void* q, void* o, const uint8_t* kv_data, const void* kv_param, const int32_t* kv_indptr,
                       const int32_t* kv_indices, const int32_t* last_page_len, int batch, int q_heads, int kv_heads,
                       int head_dim, int page_size, int layer, int n_layers, float sm_scale, flatq_stream_t stream) {
    FLATQ_CHECK(q && o && kv_data && kv_param && kv_indptr && kv_indices && last_page_len, FLATQ_ERR_INVALID, "kv_decode: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "kv_decode: head_dim=%d (only 128)", head_dim);
    FLATQ_CHECK(batch > 0 && q_heads > 0 && kv_heads > 0 && q_heads % kv_heads == 0, FLATQ_ERR_INVALID, "kv_decode: heads");
    DecodeParams p;
    p.q = reinterpret_cast<const __half*>(q); p.o = reinterpret_cast<__half*>(o);
    p.kv_data = kv_data; p.kv_param = reinterpret_cast<const __half*>(kv_param);
    p.kv_indptr = kv_indptr; p.kv_indices = kv_indices; p.last_page_len = last_page_len;
    p.batch = batch; p.q_heads = q_heads; p.kv_heads = kv_heads; p.D = head_dim; p.page_size = page_size;
    p.layer = layer; p.n_layers = n_layers; p.sm_scale = sm_scale;
    decode_attn_i4_kernel<<<dim3(batch, q_heads), 128, 0, (cudaStream_t)stream>>>(p);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

// split-KV decode attention: partial softmax states of up to kDecMaxSplits splits per (batch, query head)
static constexpr int kDecMaxSplits = 64;

size_t flatq_kv_decode_i4_split_workspace_bytes(int batch, int q_heads) {
    return align_up((size_t)(batch > 0 ? batch : 0) * (q_heads > 0 ? q_heads : 0) * kDecMaxSplits * kDecPartStride * sizeof(float), 256);
}

int flatq_kv_decode_i4_split(const void* q, void* o, const uint8_t* kv_data, const void* kv_param, const int32_t* kv_indptr,
                             const int32_t* kv_indices, const int32_t* last_page_len, int batch, int q_heads, int kv_heads,
                             int head_dim, int page_size, int layer, int n_layers, float sm_scale, int max_tokens_hint,
                             void* ws, size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(q && o && kv_data && kv_param && kv_indptr && kv_indices && last_page_len && ws, FLATQ_ERR_INVALID,
                "kv_decode_split: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "kv_decode_split: head_dim=%d (only 128)", head_dim);
    FLATQ_CHECK(batch > 0 && q_heads > 0 && kv_heads > 0 && q_heads % kv_heads == 0, FLATQ_ERR_INVALID, "kv_decode_split: heads");
    FLATQ_CHECK(page_size > 0 && n_layers > 0 && layer >= 0 && layer < n_layers, FLATQ_ERR_INVALID, "kv_decode_split: cache arguments");
    FLATQ_CHECK(aligned16(q) && aligned16(kv_data) && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, FLATQ_ERR_INVALID,
                "kv_decode_split: alignment");
    FLATQ_CHECK(ws_bytes >= flatq_kv_decode_i4_split_workspace_bytes(batch, q_heads), FLATQ_ERR_WORKSPACE,
                "kv_decode_split: workspace too small");
    const int G = q_heads / kv_heads;
    DecodeParams p;
    p.q = reinterpret_cast<const __half*>(q); p.o = reinterpret_cast<__half*>(o);
    p.kv_data = kv_data; p.kv_param = reinterpret_cast<const __half*>(kv_param);
    p.kv_indptr = kv_indptr; p.kv_indices = kv_indices; p.last_page_len = last_page_len;
    p.batch = batch; p.q_heads = q_heads; p.kv_heads = kv_heads; p.D = head_dim; p.page_size = page_size;
    p.layer = layer; p.n_layers = n_layers; p.sm_scale = sm_scale;
    float* part = reinterpret_cast<float*>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    {
        // tensor-core variant (decode_attn_mma.cuh): the query heads of a kv head are the rows of an m16n8k16 tile
        static const bool simt = getenv("FLATQ_DECODE_SIMT") != nullptr;   // A/B aid: the scalar online-softmax kernel
        if (!simt && G >= 2 && G <= 16) {   // G = 1 (no grouped queries): 15 of the tile's 16 rows would be padding; the scalar kernel wins
            // splits: fill ONE wave of resident CTAs (2 per SM: registers) -- rounding the split count up would leave a
            // second, mostly empty wave (16 x 8 heads: 3 splits = 384 CTAs on 296 slots)
            const int64_t base = (int64_t)batch * kv_heads;
            int64_t splits = (2 * (int64_t)num_sms()) / base;
            if (max_tokens_hint > 0) {   // no split shorter than one 32-token tile per warp
                const int64_t by_len = ((int64_t)max_tokens_hint + kDmWarps * kDmTile - 1) / (kDmWarps * kDmTile);
                if (splits > by_len) splits = by_len;
            }
            if (splits > kDecMaxSplits) splits = kDecMaxSplits;
            if (splits < 1) splits = 1;
            static PerDeviceOnce attr;
            if (attr.first_time())
                FLATQ_CUDA(cudaFuncSetAttribute(decode_attn_i4_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDmSmem));
            const dim3 grid((unsigned)splits, (unsigned)kv_heads, (unsigned)batch);
            FLATQ_CUDA(launch_pdl(decode_attn_i4_mma_kernel<false>, grid, dim3(kDmThreads), (size_t)kDmSmem, st, p, DecodeFront{}, (int)splits, part));
            FLATQ_CUDA(launch_pdl(decode_attn_merge_kernel, dim3(batch * q_heads), dim3(128), 0, st, (const float*)part, (int)splits,
                                  reinterpret_cast<__half*>(o)));
            return FLATQ_OK;
        }
    }
    const int gh = G % 4 == 0 ? 4 : G % 2 == 0 ? 2 : 1;
    const int64_t base = (int64_t)batch * kv_heads * (G / gh);
    // enough CTAs for two per SM, but no split shorter than 64 tokens (4 chunks of 16) when the caller bounds the length
    int64_t splits = (2 * (int64_t)num_sms() + base - 1) / base;
    if (max_tokens_hint > 0) {
        const int64_t by_len = ((int64_t)max_tokens_hint + 63) / 64;
        if (splits > by_len) splits = by_len;
    }
    if (splits > kDecMaxSplits) splits = kDecMaxSplits;
    if (splits < 1) splits = 1;
    const dim3 grid((unsigned)splits, (unsigned)(kv_heads * (G / gh)), (unsigned)batch);
    if (gh == 4) FLATQ_CUDA(launch_pdl(decode_attn_i4_split_kernel<4>, grid, dim3(kDecSplitThreads), 0, st, p, (int)splits, part));
    else if (gh == 2) FLATQ_CUDA(launch_pdl(decode_attn_i4_split_kernel<2>, grid, dim3(kDecSplitThreads), 0, st, p, (int)splits, part));
    else FLATQ_CUDA(launch_pdl(decode_attn_i4_split_kernel<1>, grid, dim3(kDecSplitThreads), 0, st, p, (int)splits, part));
    FLATQ_CUDA(launch_pdl(decode_attn_merge_kernel, dim3(batch * q_heads), dim3(128), 0, st, (const float*)part, (int)splits,
                          reinterpret_cast<__half*>(o)));
    return FLATQ_OK;
}

int flatq_kv_decode_fused(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_k, int64_t ld_v, const void* cos_t,
                          const void* sin_t, const int32_t* pos, const void* t_q, const void* t_k, const int32_t* slot, void* o,
                          uint8_t* kv_data, void* kv_param, const int32_t* kv_indptr, const int32_t* kv_indices,
                          const int32_t* last_page_len, int batch, int q_heads, int kv_heads, int head_dim, int page_size, int layer,
                          int n_layers, float sm_scale, int max_tokens_hint, void* ws, size_t ws_bytes, flatq_stream_t stream) {
    FLATQ_CHECK(q && k && v && cos_t && sin_t && pos && t_q && t_k && slot && o && kv_data && kv_param && kv_indptr && kv_indices &&
                last_page_len && ws, FLATQ_ERR_INVALID, "kv_decode_fused: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "kv_decode_fused: head_dim=%d (only 128)", head_dim);
    FLATQ_CHECK(batch > 0 && q_heads > 0 && kv_heads > 0 && q_heads % kv_heads == 0, FLATQ_ERR_INVALID, "kv_decode_fused: heads");
    const int G = q_heads / kv_heads;
    FLATQ_CHECK(G >= 2 && G <= 16, FLATQ_ERR_UNSUPPORTED,
                "kv_decode_fused: %d query heads per kv head (2..16: the grouped-query tensor-core kernel); use flatq_kv_fused + flatq_kv_decode_i4_split", G);
    FLATQ_CHECK(page_size > 0 && n_layers > 0 && layer >= 0 && layer < n_layers, FLATQ_ERR_INVALID, "kv_decode_fused: cache arguments");
    FLATQ_CHECK(ld_q % 4 == 0 && ld_k % 4 == 0 && ld_v % 4 == 0 && ld_q >= (int64_t)q_heads * 128 && ld_k >= (int64_t)kv_heads * 128 &&
                ld_v >= (int64_t)kv_heads * 128, FLATQ_ERR_INVALID, "kv_decode_fused: row strides");
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    FLATQ_CHECK(al8(q) && al8(k) && al8(v) && al8(cos_t) && al8(sin_t) && aligned16(t_q) && aligned16(t_k) && aligned16(kv_data) &&
                (reinterpret_cast<uintptr_t>(ws) & 15) == 0, FLATQ_ERR_INVALID, "kv_decode_fused: alignment");
    FLATQ_CHECK(ws_bytes >= flatq_kv_decode_i4_split_workspace_bytes(batch, q_heads), FLATQ_ERR_WORKSPACE, "kv_decode_fused: workspace too small");
    DecodeParams p;
    p.q = nullptr; p.o = reinterpret_cast<__half*>(o);
    p.kv_data = kv_data; p.kv_param = reinterpret_cast<const __half*>(kv_param);
    p.kv_indptr = kv_indptr; p.kv_indices = kv_indices; p.last_page_len = last_page_len;
    p.batch = batch; p.q_heads = q_heads; p.kv_heads = kv_heads; p.D = head_dim; p.page_size = page_size;
    p.layer = layer; p.n_layers = n_layers; p.sm_scale = sm_scale;
    DecodeFront f;
    f.q = reinterpret_cast<const __half*>(q); f.k = reinterpret_cast<const __half*>(k); f.v = reinterpret_cast<const __half*>(v);
    f.ld_q = ld_q; f.ld_k = ld_k; f.ld_v = ld_v;
    f.cos_t = reinterpret_cast<const __half*>(cos_t); f.sin_t = reinterpret_cast<const __half*>(sin_t); f.pos = pos;
    f.t_q = reinterpret_cast<const __half*>(t_q); f.t_k = reinterpret_cast<const __half*>(t_k); f.slot = slot;
    f.kv_data = kv_data; f.kv_param = reinterpret_cast<__half*>(kv_param);
    float* part = reinterpret_cast<float*>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t base = (int64_t)batch * kv_heads;
    int64_t splits = (2 * (int64_t)num_sms()) / base;   // one wave of resident CTAs (see flatq_kv_decode_i4_split)
    if (max_tokens_hint > 0) {
        const int64_t by_len = ((int64_t)max_tokens_hint + kDmWarps * kDmTile - 1) / (kDmWarps * kDmTile);
        if (splits > by_len) splits = by_len;
    }
    if (splits > kDecMaxSplits) splits = kDecMaxSplits;
    if (splits < 1) splits = 1;
    static PerDeviceOnce attr;
    if (attr.first_time())
        FLATQ_CUDA(cudaFuncSetAttribute(decode_attn_i4_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDmSmem));
    const dim3 grid((unsigned)splits, (unsigned)kv_heads, (unsigned)batch);
    FLATQ_CUDA(launch_pdl(decode_attn_i4_mma_kernel<true>, grid, dim3(kDmThreads), (size_t)kDmSmem, st, p, f, (int)splits, part));
    FLATQ_CUDA(launch_pdl(decode_attn_merge_kernel, dim3(batch * q_heads), dim3(128), 0, st, (const float*)part, (int)splits,
                          reinterpret_cast<__half*>(o)));
    return FLATQ_OK;
}

int flatq_rmsnorm(const void* x, const void* weight, int64_t rows, int64_t K, float mean_dim, float eps, void* y,
                  flatq_stream_t stream) {
    FLATQ_CHECK(x && y, FLATQ_ERR_INVALID, "rmsnorm: null pointer");
    FLATQ_CHECK(rows > 0 && K > 0 && K % 8 == 0 && rows < (1ll << 31) && K < (1ll << 31), FLATQ_ERR_INVALID,
                "rmsnorm: rows=%lld, K=%lld (K must be a multiple of 8)", (long long)rows, (long long)K);
    FLATQ_CHECK(mean_dim > 0.f, FLATQ_ERR_INVALID, "rmsnorm: mean_dim");
    FLATQ_CHECK(aligned16(x) && aligned16(y) && (!weight || aligned16(weight)), FLATQ_ERR_INVALID, "rmsnorm: alignment");
    const int64_t cap = (int64_t)num_sms() * 8;
    FLATQ_CUDA(launch_pdl(rmsnorm_f16_kernel, dim3((unsigned)(rows < cap ? rows : cap)), dim3(kNormThreads), 0, (cudaStream_t)stream,
                          reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(weight), reinterpret_cast<__half*>(y),
                          (int)rows, (int)K, mean_dim, eps, (const __half*)nullptr, (__half*)nullptr));
    return FLATQ_OK;
}

int flatq_add_rmsnorm(const void* h, const void* y_add, const void* weight, int64_t rows, int64_t K, float mean_dim, float eps,
                      void* h_out, void* normed, flatq_stream_t stream) {
    FLATQ_CHECK(h && y_add && h_out && normed, FLATQ_ERR_INVALID, "add_rmsnorm: null pointer");
    FLATQ_CHECK(rows > 0 && K > 0 && K % 8 == 0 && rows < (1ll << 31) && K < (1ll << 31), FLATQ_ERR_INVALID,
                "add_rmsnorm: rows=%lld, K=%lld (K must be a multiple of 8)", (long long)rows, (long long)K);
    FLATQ_CHECK(mean_dim > 0.f, FLATQ_ERR_INVALID, "add_rmsnorm: mean_dim");
    FLATQ_CHECK(aligned16(h) && aligned16(y_add) && aligned16(h_out) && aligned16(normed) && (!weight || aligned16(weight)),
                FLATQ_ERR_INVALID, "add_rmsnorm: alignment");
    const int64_t cap = (int64_t)num_sms() * 8;
    FLATQ_CUDA(launch_pdl(rmsnorm_f16_kernel, dim3((unsigned)(rows < cap ? rows : cap)), dim3(kNormThreads), 0, (cudaStream_t)stream,
                          reinterpret_cast<const __half*>(h), reinterpret_cast<const __half*>(weight), reinterpret_cast<__half*>(normed),
                          (int)rows, (int)K, mean_dim, eps, reinterpret_cast<const __half*>(y_add), reinterpret_cast<__half*>(h_out)));
    return FLATQ_OK;
}

int flatq_rope(void* x, const void* cos_t, const void* sin_t, int64_t tokens, int heads, int head_dim, flatq_stream_t stream) {
    FLATQ_CHECK(x && cos_t && sin_t, FLATQ_ERR_INVALID, "rope: null pointer");
    FLATQ_CHECK(head_dim == 128, FLATQ_ERR_UNSUPPORTED, "rope: head_dim=%d (only 128)", head_dim);
    FLATQ_CHECK(tokens > 0 && heads > 0 && tokens * heads < (1ll << 31), FLATQ_ERR_INVALID, "rope: tokens / heads");
    FLATQ_CHECK(aligned16(x) && aligned16(cos_t) && aligned16(sin_t), FLATQ_ERR_INVALID, "rope: alignment");
    FLATQ_CUDA(launch_pdl(rope_f16_kernel, dim3(grid_for((size_t)tokens * heads * 8, 256, 8)), dim3(256), 0, (cudaStream_t)stream,
                          reinterpret_cast<__half*>(x), reinterpret_cast<const __half*>(cos_t), reinterpret_cast<const __half*>(sin_t),
                          (int)tokens, heads));
    return FLATQ_OK;
}

int flatq_silu_mul(const void* gu, int64_t rows, int64_t inter, void* out, flatq_stream_t stream) {
    FLATQ_CHECK(gu && out, FLATQ_ERR_INVALID, "silu_mul: null pointer");
    FLATQ_CHECK(rows > 0 && inter > 0 && inter % 8 == 0 && rows < (1ll << 31) && inter < (1ll << 30), FLATQ_ERR_INVALID,
                "silu_mul: rows=%lld, inter=%lld (inter must be a multiple of 8)", (long long)rows, (long long)inter);
    FLATQ_CHECK(aligned16(gu) && aligned16(out), FLATQ_ERR_INVALID, "silu_mul: alignment");
    FLATQ_CUDA(launch_pdl(silu_mul_f16_kernel, dim3(grid_for((size_t)rows * (inter / 8), 256, 8)), dim3(256), 0, (cudaStream_t)stream,
                          reinterpret_cast<const __half*>(gu), reinterpret_cast<__half*>(out), (int)rows, (int)inter));
    return FLATQ_OK;
}

// Debug only (not part of the public header): select the GEMM variants listed at opt_gemm_1sm(); -1 = from the environment
// Debug only: CTA 0 of the one-launch decode FlatLinear dumps the activation codes [M, K] (int8, GEMM operand order) and
// scales [M] it computed (tests check the result exactly against the integer arithmetic on those codes); null = off
void flatq_debug_set_decode_dump(int8_t* codes, float* sx) { g_decode_dump_codes = codes; g_decode_dump_sx = sx; }

void flatq_debug_set_gemm_options(int one_sm, int streamk) { g_gemm_1sm = one_sm; g_streamk = streamk; }

// Debug only (not part of the public header): role-level cycle counters of the GEMM, TN=128 fp16 epilogue.
// prof: int64 [grid, kGemmWarps, 5]
int flatq_debug_gemm_profile(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, int64_t M,
                             int64_t N, int64_t K, void* y, long long* prof, int dbg, flatq_stream_t stream) {
    CUtensorMap tw, tx;
    int rc = make_map_u8(&tw, w_packed, N, K / 2, K / 2, kGemmBlockK / 2, kGemmBlockW, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_u8(&tx, xq8, M, K, K, 128, 256, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    GemmParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.acc_shift = 4;
    p.out = y; p.s_x = s_x; p.s_w = s_w; p.bias = nullptr; p.prof = prof; p.dbg = dbg; p.k_splits = 1;
    p.streamk = 0; p.sk_part = nullptr; p.sk_flags = nullptr; p.epoch = 0;
    p.grp_off = nullptr; p.grp_tile = nullptr; p.n_groups = 0; p.t_upper = 0; p.rows_per_owner = 0; p.tp_rank = 0; p.silu_pairs = 0;
    auto kern = gemm_w4a8_kernel<256, kEpiF16, true>;
    constexpr int smem = GemmSmem<256>::kBytes;
    FLATQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int tiles = (int)(((N + 127) / 128) * ((M + 255) / 256));
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(tw, tx, p);
    FLATQ_CUDA(cudaGetLastError());
    return FLATQ_OK;
}

}  // extern "C"
