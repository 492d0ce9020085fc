/* flatq_b200 -- C ABI of the B200 (sm_100a) FlatQuant W4A4 FlatLinear hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 * cites the reference interface it replaces (paths relative to the FlatQuant repository).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; tensors are dense row-major
 *   - `stream` is a cudaStream_t (0 = legacy default stream, which is what the reference uses:
 *     deploy/kernels/quant.cu:62,97 and gemm.cu:41 launch on it)
 *   - return value 0 on success, negative flatq_status on failure; flatq_last_error() returns a
 *     thread-local message (the reference throws std::runtime_error: deploy/kernels/include/common.h:16-23;
 *     the Python shim turns a non-zero status into RuntimeError, so callers see the same behaviour)
 *   - nothing is allocated inside; outputs and workspaces are caller-owned
 *     (the reference allocates outputs in its pybind layer, bindings.cpp:20,39,74 -- the Python shim does that)
 *   - 4-bit storage: two's-complement nibbles, element 2j in bits [3:0] and 2j+1 in bits [7:4] of byte j
 *     (deploy/functional/quantization.py:49-56, deploy/kernels/include/int4.h:86-98)
 *   - thread safe: no mutable global state except the lazily resolved driver entry point
 */
#ifndef FLATQ_B200_H_
#define FLATQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLATQ_ABI_VERSION 1

typedef void* flatq_stream_t; /* cudaStream_t */

enum flatq_status {
    FLATQ_OK = 0,
    FLATQ_ERR_INVALID = -1,     /* bad shape / alignment / null pointer        */
    FLATQ_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement    */
    FLATQ_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed         */
    FLATQ_ERR_WORKSPACE = -4    /* workspace too small                         */
};

enum flatq_dtype { FLATQ_F16 = 0, FLATQ_BF16 = 1, FLATQ_I32 = 2 };

int flatq_abi_version(void);
const char* flatq_last_error(void);
/* 1 when the current device is compute capability 10.x (tcgen05 present), else 0 */
int flatq_device_supported(void);

/* ---- a8: nibble pack / unpack (deploy/functional/quantization.py:49-82) -------------------- */
/* q: int8 [rows, cols] in [-8,7], cols % 16 == 0  ->  packed uint8 [rows, cols/2] */
int flatq_pack_i4(const int8_t* q, uint8_t* packed, int64_t rows, int64_t cols, flatq_stream_t stream);
/* packed uint8 [rows, cols_packed], cols_packed % 4 == 0  ->  int32 [rows, 2*cols_packed] */
int flatq_unpack_i4(const uint8_t* packed, int32_t* out, int64_t rows, int64_t cols_packed, flatq_stream_t stream);

/* ---- a10: deploy._CUDA.sym_quant -> sym_quant_host (deploy/kernels/include/quant.h:6-13,
 *      quant.cu:13-63).  x fp16 [rows, cols], scale fp16 [rows] -> packed [rows, cols/2];
 *      q = clamp(rn(half(x/scale)), -8, 7).  cols % 8 == 0. */
int flatq_sym_quant(const void* x_f16, const void* scale_f16, int64_t rows, int64_t cols, uint8_t* q_packed,
                    flatq_stream_t stream);

/* ---- a12: deploy._CUDA.sym_dequant -> sym_dequant_host (quant.h:16-23, quant.cu:66-101) plus the
 *      bias add of deploy/nn/linear.py:51-54.  acc int32 [rows, cols], scale_row fp16 [rows],
 *      scale_col fp16 [cols], bias (out dtype, may be NULL) -> y [rows, cols].
 *      mode 0: y = acc*s_row*s_col (+bias) in fp32, one rounding (exact form; matches the fake-quant oracle)
 *      mode 1: bit-level emulation of the reference kernel (half(int(acc/10))*10 in half arithmetic; fp16 only) */
int flatq_sym_dequant(const int32_t* acc, const void* scale_row_f16, const void* scale_col_f16, const void* bias,
                      int64_t rows, int64_t cols, void* y, int out_dtype, int mode, flatq_stream_t stream);

/* ---- a11: deploy._CUDA.matmul -> matmul_host (deploy/kernels/include/gemm.h:6-13, gemm.cu:7-46)
 *      C int32 [M, N] = A[M, K] . B[N, K]^T, both packed 4-bit ([M, K/2], [N, K/2]); K % 64 == 0.
 *      workspace: flatq_gemm_i4_workspace_bytes(M, K) bytes (int8 expansion of A). */
size_t flatq_gemm_i4_workspace_bytes(int64_t M, int64_t K);
int flatq_gemm_i4(const uint8_t* A_packed, const uint8_t* B_packed, int64_t M, int64_t N, int64_t K, int32_t* C,
                  void* workspace, size_t workspace_bytes, flatq_stream_t stream);

/* ---- a11 + a12 fused: deploy.nn.Linear4bit.forward (deploy/nn/linear.py:40-54) in one kernel.
 *      y[M, N] = (A . B^T) * s_a[m] * s_b[n] (+ bias[n]);  scales fp16, y/bias out_dtype. */
int flatq_linear4bit(const uint8_t* A_packed, const void* s_a_f16, const uint8_t* B_packed, const void* s_b_f16,
                     const void* bias, int64_t M, int64_t N, int64_t K, void* y, int out_dtype, void* workspace,
                     size_t workspace_bytes, flatq_stream_t stream);
size_t flatq_linear4bit_workspace_bytes(int64_t M, int64_t N, int64_t K);

/* ---- a2/a3: kronecker_matmul (flatquant/flat_utils.py:6-17) / *TransMatrix eval forward
 *      (flatquant/trans_utils.py:21-25,97-101).  z[t] = L^T . reshape(x[t], n1, n2) . R, fp16.
 *      left [n1,n1] or NULL (identity), right [n2,n2] or NULL; n2 % 16 == 0, n1,n2 <= 256. */
int flatq_kron_transform(const void* x_f16, const void* left_f16, const void* right_f16, int64_t tokens, int n1,
                         int n2, void* z_f16, flatq_stream_t stream);

/* ---- a9: deploy.nn.OnlineTrans + Quantizer -> kron_matmul / block_matmul
 *      (deploy/functional/online_trans.py:113-141, deploy/kernels/kron_matmul.py:192-266).
 *      transform, per-token scale (clip factors are the *multiplicative* factors, i.e. already
 *      sigmoid(clip_factor_a_*) or the clip ratio), quantise, pack.
 *      outputs (each may be NULL): xq_packed [tokens, K/2], scales_f16 [tokens], scales_f32 [tokens]. */
int flatq_transform_quant(const void* x_f16, const void* left_f16, const void* right_f16, int64_t tokens, int n1,
                          int n2, float clip_max, float clip_min, uint8_t* xq_packed, void* scales_f16,
                          float* scales_f32, flatq_stream_t stream);

/* ---- a10 (fused form): deploy.nn.Quantizer.forward (deploy/nn/quantization.py:13-36): per-token
 *      absmax scale + quantise + pack in one pass.  half_math != 0 reproduces the fp16 arithmetic of the
 *      deploy path (fp16 scale, __hdiv); 0 uses the fp32 arithmetic of the fake-quant LAC path. */
int flatq_quantize_rows(const void* x_f16, int64_t tokens, int64_t K, float clip_max, float clip_min, int half_math,
                        uint8_t* xq_packed, void* scales_f16, float* scales_f32, flatq_stream_t stream);

/* ---- the whole FlatLinear: a2+a4+a11+a12 (FlatQuantizedLinear eval forward after the caller's
 *      qa_trans: flatquant/flat_linear.py:75-80 + trans_utils.py:97-101;
 *      == OnlineTrans -> Quantizer -> Linear4bit of benchmarks/qlinear_benchmark.py:70-74).
 *      x fp16 [M, K=n1*n2]; left/right may be NULL (no transform / single-sided);
 *      w_packed [N, K/2]; s_w fp32 [N]; bias out_dtype [N] or NULL; y [M, N] out_dtype.
 *      workspace >= flatq_linear_fused_workspace_bytes(M, N, K).
 *      One or two tokens (a decode step) with factor sizes n1 % 16 == 0, n2 in {64, 128} run as ONE launch: every CTA of
 *      the weight-stream GEMV computes the transform + quantisation of the token(s) itself while its first weight rows
 *      are in flight (csrc/flatlinear_decode.cuh); other shapes are the transform kernel followed by the GEMM. */
size_t flatq_linear_fused_workspace_bytes(int64_t M, int64_t N, int64_t K);
int flatq_linear_fused(const void* x_f16, const void* left_f16, const void* right_f16, int n1, int n2,
                       float clip_max, float clip_min, const uint8_t* w_packed, const float* s_w, const void* bias,
                       int64_t M, int64_t N, void* y, int out_dtype, void* workspace, size_t workspace_bytes,
                       flatq_stream_t stream);

/* ---- the two halves of flatq_linear_fused, exposed so one quantised activation can feed several
 *      weight matrices (q/k/v, gate/up: the reference re-runs OnlineTrans+Quantizer three times on the
 *      same hidden state, deploy/transformers/modeling_llama.py:66-78).
 *      flatq_act_quant: transform + per-token quantise into the GEMM's activation operand format
 *        xq8 int8 [M, K] (codes; K order permuted [0,2,4,6,1,3,5,7] inside each group of 8 -- opaque),
 *        s_x fp32 [M].
 *      flatq_gemm_w4a8: y[M,N] = (xq8 . W^T) * s_x[m] * s_w[n] (+bias), W packed [N, K/2]; K % 32 == 0. */
int flatq_act_quant(const void* x_f16, const void* left_f16, const void* right_f16, int n1, int n2, float clip_max,
                    float clip_min, int64_t M, int8_t* xq8, float* s_x, flatq_stream_t stream);
/*      M <= 2: dp4a GEMV; 3 <= M <= 32 (K % 128 == 0): the weight-streaming mma.sync kernel -- neither needs a workspace.
 *      workspace (may be NULL): with flatq_gemm_w4a8_workspace_bytes(M, N, K) bytes, 33 <= M <= 64 split K
 *      across CTAs (int32 atomics + a finalize pass) so that the packed-weight stream uses the whole chip, and
 *      prefill shapes whose last wave of output tiles is partial deal that wave's K ranges to all SMs (stream-K:
 *      raw int32 partials + per-launch epoch flags in the workspace; needs no zeroing, must not be shared by
 *      launches that may run concurrently).  Results are bit-identical with and without the workspace. */
size_t flatq_gemm_w4a8_workspace_bytes(int64_t M, int64_t N, int64_t K);
int flatq_gemm_w4a8(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, const void* bias,
                    int64_t M, int64_t N, int64_t K, void* y, int out_dtype, void* workspace, size_t workspace_bytes,
                    flatq_stream_t stream);

/* ---- a6 + the MLP activation: the fused [gate; up] FlatLinear followed by act_fn(gate) * up
 *      (transformers' LlamaMLP.forward under flatquant/model_tools/llama_utils.py:16-54: down_proj(act_fn(gate_proj(x))
 *      * up_proj(x)); the reference runs two F.linear calls, F.silu and a multiply).
 *      w_packed [N, K/2] / s_w [N] hold the two projections ROW-INTERLEAVED: row 2j = gate_proj row j, row 2j + 1 =
 *      up_proj row j (N = 2 * intermediate_size, N % 16 == 0; no bias: Llama's MLP has none).
 *      out [M, N/2] fp16 = half(silu_f32(g)) * u with g, u the fp16-rounded GEMM outputs -- bit-identical to
 *      flatq_gemm_w4a8 into [M, N] followed by flatq_silu_mul on the de-interleaved columns.  For M > 128 the
 *      activation runs in the GEMM epilogue (the [M, N] pre-activation never reaches HBM) and no workspace is needed;
 *      smaller M use the workspace (flatq_gemm_w4a8_silu_workspace_bytes) for the pre-activation + split-K. */
size_t flatq_gemm_w4a8_silu_workspace_bytes(int64_t M, int64_t N, int64_t K);
int flatq_gemm_w4a8_silu(const int8_t* xq8, const float* s_x, const uint8_t* w_packed_il, const float* s_w_il, int64_t M,
                         int64_t N, int64_t K, void* out_f16, void* workspace, size_t workspace_bytes,
                         flatq_stream_t stream);

/* ---- the whole MLP front half: transform -> quantise -> fused [gate; up] GEMM -> act_fn(gate) * up, i.e.
 *      flatq_act_quant + flatq_gemm_w4a8_silu (same arguments, same row-interleaved weight); one launch for one or two
 *      tokens (see flatq_linear_fused).  out [M, N/2] fp16. */
size_t flatq_linear_fused_silu_workspace_bytes(int64_t M, int64_t N, int64_t K);
int flatq_linear_fused_silu(const void* x_f16, const void* left_f16, const void* right_f16, int n1, int n2, float clip_max,
                            float clip_min, const uint8_t* w_packed_il, const float* s_w_il, int64_t M, int64_t N,
                            void* out_f16, void* workspace, size_t workspace_bytes, flatq_stream_t stream);

/* ---- a6 for mixture-of-experts layers: the routed-expert FlatLinears of the DeepSeek block (BASELINE configs[4];
 *      deepseek_v3/model.py:677-683 -- a Python loop over experts, each `expert(x[idx])` -- under
 *      flatquant/model_tools/deepseekv3_utils.py:413-470).  ONE launch for all experts of a stage:
 *        rows [group_offsets[e], group_offsets[e+1]) of xq8 / s_x / y belong to expert e  (tokens sorted by expert),
 *        w_packed [n_groups, N, K/2], s_w [n_groups, N], bias [n_groups, N] or NULL,
 *        y[m, :] = (xq8[m] . W_e^T) * s_x[m] * s_w[e, :] (+ bias[e, :])   -- bit-identical to n_groups flatq_gemm_w4a8 calls.
 *      group_offsets is a DEVICE array (int32, n_groups + 1, non-decreasing, [0] == 0, [n_groups] <= M): the host does
 *      not read the routing result.  rows_per_group_hint picks the token tile (<= 0: M / n_groups).
 *      workspace >= flatq_gemm_w4a8_grouped_workspace_bytes(n_groups), 4-byte aligned. */
size_t flatq_gemm_w4a8_grouped_workspace_bytes(int n_groups);
int flatq_gemm_w4a8_grouped(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w,
                            const void* bias, const int32_t* group_offsets, int n_groups, int64_t M, int64_t N, int64_t K,
                            int rows_per_group_hint, void* y, int out_dtype, void* workspace, size_t workspace_bytes,
                            flatq_stream_t stream);

/* ---- (e) tensor parallel: row-parallel linears with the reduction done by THIS library over NVLink peer memory ----------
 * The reference sums the ranks' partial products with dist.all_reduce after F.linear (flatquant/model_tools/
 * deepseekv3_utils.py:98-105; vllm_custom/.../llama_flatquant.py:117-126,261-289).  Here the caller hands over mapped peer
 * memory -- from NCCL's window registration (ncclCommWindowRegister + ncclGetLsaPointer / the multimem pointer of an
 * ncclComm_t), cuMemMap of fabric handles, or torch.distributed._symmetric_memory as flatquant_b200/peer_reduce.py does --
 * and the kernels move the bytes: see csrc/peer_reduce.cuh.
 *   flatq_tp_ctx: rank / world (2..8), signal_pads[r] = rank r's flag area (zero-initialised symmetric memory of
 *     signal_pad_bytes >= 2 * 96 * (world + 1) * 4; used only by these calls, never reset).
 *   flatq_gemm_w4a8_rowparallel: the prefill GEMM whose epilogue IS the reduce-scatter: rank `rank` stores every 256-token
 *     output tile into slot `rank` of the staging buffer [world][M / world][N] fp16 of the tile's owner
 *     (peer_staging[owner]; rows [o M/world, (o+1) M/world) belong to rank o).  M % (256 world) == 0.
 *   flatq_peer_reduce_bcast: owner-side sum of the world slots in fp32 (+ bias), rounded to fp16, written to rows
 *     [rank M/world, ...) of EVERY rank's [M, N] output through out_peers[r] (unicast peer stores: the receive side bounds
 *     an all-gather and a multicast store would also send this rank's copy back to it); out_mc (multimem.st) is used
 *     only when out_peers == NULL.  Begins and ends with a barrier over the signal pads.
 *   flatq_peer_allreduce: in-switch all-reduce (multimem.ld_reduce / multimem.st) of a symmetric fp16 [rows, N] buffer, for
 *     the shapes the fused path does not take (decode sizes); src_mc / dst_mc are multicast addresses (may be equal). */
typedef struct flatq_tp_ctx {
    int rank, world;
    void* signal_pads[8];
    size_t signal_pad_bytes;
} flatq_tp_ctx;
int flatq_gemm_w4a8_rowparallel(const int8_t* xq8, const float* s_x, const uint8_t* w_packed, const float* s_w, int64_t M,
                                int64_t N, int64_t K, const flatq_tp_ctx* tp, void* const* peer_staging, flatq_stream_t stream);
int flatq_peer_reduce_bcast(const flatq_tp_ctx* tp, const void* staging, void* out_mc, void* const* out_peers, const void* bias_f16,
                            int64_t M, int64_t N, flatq_stream_t stream);
int flatq_peer_allreduce(const flatq_tp_ctx* tp, const void* src_mc, void* dst_mc, const void* bias_f16, int64_t rows, int64_t N,
                         flatq_stream_t stream);

/* ---- a14: qattention side path ------------------------------------------------------------------
 * flatq_hadamard: out = scale * x . H_n per row, n a power of two in [8, 32768], fp16 in/out, fp32 butterflies.
 *   Replaces fast_hadamard_transform.hadamard_transform (third-party/fast-hadamard-transform/csrc/
 *   fast_hadamard_transform.cpp:71-112), as used by deploy/functional/online_trans.py:144-151 and
 *   deploy/transformers/kv_cache.py:64-67.
 * flatq_kv_quant: per (token, head) asymmetric 4-bit quantise + pack of K or V (deploy/transformers/kv_cache.py:11-51,
 *   non-LAC branch) and, when kv_data != NULL, the scatter into the paged cache (init_kv_i4 / append_kv_i4:
 *   deploy/kernels/bindings.cpp:118-245, include/flashinfer/page.cuh:119-219).  head_dim == 128.
 *   slot[token] = page_id * page_size + offset.  Cache layout: data [pages, layers, 2, heads, page_size, 64] u8,
 *   param [pages, layers, 2, heads, page_size, 2] f16 (scale, zero)  (kv_cache.py:180-193).
 * flatq_kv_decode_i4: single-query attention over the packed cache (batch_decode_i4: bindings.cpp:118-170,
 *   include/flashinfer/decode.cuh:480-689): o = softmax(q.K^T * sm_scale) . V with x_hat = nibble*scale - zero. */
int flatq_hadamard(const void* x_f16, void* out_f16, int64_t rows, int n, float scale, flatq_stream_t stream);
/* flatq_hadamard_kn: matmul_hadU_cuda for n = K * 2^m (deploy/functional/online_trans.py:144-151: hadamard_transform over
 *   the 2^m axis of x viewed as [rows, K, n/K], scaled, rounded to fp16, then `hadK @ input` over the K axis; the fused
 *   counterparts in the third-party library are fast_hadamard_transform_cuda.cu:37-127,276-441).  hadK: fp16 [K, K]
 *   row-major (the reference's table for K, see flatquant_b200/deploy/functional/data), K % 4 == 0, K <= 172. */
int flatq_hadamard_kn(const void* x_f16, void* out_f16, const void* hadK_f16, int64_t rows, int n, int K, float scale,
                      flatq_stream_t stream);
int flatq_kv_quant(const void* x_f16, int64_t tokens, int heads, int head_dim, uint8_t* q_out, void* param_out_f16,
                   uint8_t* kv_data, void* kv_param_f16, const int32_t* slot, int layer, int n_layers, int kv_select,
                   int page_size, flatq_stream_t stream);
int flatq_kv_decode_i4(const void* q_f16, void* o_f16, const uint8_t* kv_data, const void* kv_param_f16,
                       const int32_t* kv_indptr, const int32_t* kv_indices, const int32_t* last_page_len, int batch,
                       int q_heads, int kv_heads, int head_dim, int page_size, int layer, int n_layers, float sm_scale,
                       flatq_stream_t stream);

/* flatq_kv_fused: everything between the q/k/v FlatLinears and the attention in ONE launch, reading
 *   the rows (token, head) of the fused qkv GEMM output in place (row of X at X + token * ld_X + head * 128 halves):
 *     q: RoPE -> . t_q (T_k^-T) -> q_out [tokens, q_heads, 128];   k: RoPE -> . t_k -> (k_out) -> asym 4-bit -> page scatter;
 *     v: asym 4-bit -> page scatter.
 *   Replaces apply_rotary_pos_emb (deploy/transformers/modeling_llama.py:125-137), the K / Q transform, asym_quantize_and_
 *   pack_i4 and init_kv_i4 / append_kv_i4 (deploy/transformers/kv_cache.py:11-51,263-353; deploy/kernels/bindings.cpp:
 *   171-245).  cos/sin: fp16 tables [positions, 128] (NULL: no RoPE), pos[token] the row to use; t_q / t_k: fp16 [128, 128]
 *   row-major or NULL (identity); k_out: NULL at decode.  Same bits as flatq_rope + flatq_kron_transform + flatq_kv_quant. */
int flatq_kv_fused(const void* q_f16, const void* k_f16, const void* v_f16, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                   int64_t tokens, int q_heads, int kv_heads, int head_dim, const void* cos_f16, const void* sin_f16,
                   const int32_t* pos, const void* t_q_f16, const void* t_k_f16, void* q_out_f16, void* k_out_f16,
                   uint8_t* kv_data, void* kv_param_f16, const int32_t* slot, int layer, int n_layers, int page_size,
                   flatq_stream_t stream);

/* flatq_kv_decode_fused: a decode step's flatq_kv_fused + flatq_kv_decode_i4_split in one kernel + the split merge (one
 *   token per sequence; grouped-query models with 2..16 query heads per kv head, head_dim 128, both transforms and RoPE
 *   present -- other cases: the two calls).  The attention kernel's prologue rotates and transforms its group's query rows
 *   and, in the CTA whose split holds the newest token, transforms / quantises / stores that token's K and V rows into the
 *   cache slot slot[b] before attending over [0, length) (csrc/decode_attn_mma.cuh).  q / k / v: the new token's pre-RoPE rows
 *   of sequence b at X + b * ld_X + head * 128; pos[b]: RoPE table row; last_page_len[b] already counts the new token.
 *   Bit-identical (cache contents and q_t; the attention arithmetic is the same kernel) to the two calls.
 *   workspace >= flatq_kv_decode_i4_split_workspace_bytes(batch, q_heads). */
int flatq_kv_decode_fused(const void* q_f16, const void* k_f16, const void* v_f16, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                          const void* cos_f16, const void* sin_f16, const int32_t* pos, const void* t_q_f16, const void* t_k_f16,
                          const int32_t* slot, void* o_f16, uint8_t* kv_data, void* kv_param_f16, const int32_t* kv_indptr,
                          const int32_t* kv_indices, const int32_t* last_page_len, int batch, int q_heads, int kv_heads,
                          int head_dim, int page_size, int layer, int n_layers, float sm_scale, int max_tokens_hint,
                          void* workspace, size_t workspace_bytes, flatq_stream_t stream);

/* flatq_kv_decode_i4_split: the same attention with the sequence split over CTAs (flash-decoding) and one CTA per kv head
 *   serving up to four of its query heads (grouped-query attention reads each K/V row once per group): fills the chip at
 *   small batch, where flatq_kv_decode_i4 runs batch x q_heads CTAs.  max_tokens_hint (host-side upper bound of the
 *   sequence lengths, <= 0: unknown) limits the number of splits; the lengths themselves are read on the device.
 *   workspace >= flatq_kv_decode_i4_split_workspace_bytes(batch, q_heads), 16-byte aligned; results equal
 *   flatq_kv_decode_i4's up to fp32 summation order.  Opt-in until its first run on hardware. */
size_t flatq_kv_decode_i4_split_workspace_bytes(int batch, int q_heads);
int flatq_kv_decode_i4_split(const void* q_f16, void* o_f16, const uint8_t* kv_data, const void* kv_param_f16,
                             const int32_t* kv_indptr, const int32_t* kv_indices, const int32_t* last_page_len, int batch,
                             int q_heads, int kv_heads, int head_dim, int page_size, int layer, int n_layers,
                             float sm_scale, int max_tokens_hint, void* workspace, size_t workspace_bytes,
                             flatq_stream_t stream);

/* ---- elementwise glue between the FlatLinears of a decoder layer (the callers either side of the path; HF
 *      transformers' torch expressions in the reference).  One pass each, fp16 in/out, the rounding points of the
 *      expression replaced.  Opt-in in the model stack until their first run on hardware.
 * flatq_rmsnorm:  y = weight * half(float(x) * rsqrt(sum(x^2) / mean_dim + eps))   (LlamaRMSNorm.forward; weight NULL =
 *                 the weight-less norm of deploy/nn/normalization.py:4-23).  x, y [rows, K], K % 8 == 0.
 * flatq_rope:     in place, x [tokens, heads, 128], cos / sin [tokens, 128]: x * cos + rotate_half(x) * sin with every
 *                 op rounded to fp16 (apply_rotary_pos_emb of HF Llama on fp16 tensors).
 * flatq_silu_mul: gu [rows, 2 * inter] = [gate | up] (the fused gate/up GEMM output) -> out [rows, inter] =
 *                 half(silu_f32(gate)) * up (LlamaMLP.forward: act_fn(gate_proj(x)) * up_proj(x)). */
int flatq_rmsnorm(const void* x_f16, const void* weight_f16, int64_t rows, int64_t K, float mean_dim, float eps,
                  void* y_f16, flatq_stream_t stream);
/* flatq_add_rmsnorm: the residual add of a decoder layer fused into the norm that follows it: h_out = half(h + y_add)
 * (torch's fp16 add), normed = flatq_rmsnorm(h_out).  h_out may alias h.  Replaces `h = h + sublayer(...)` followed by
 * LlamaRMSNorm.forward (HF modeling_llama.py; called from deploy/transformers/modeling_llama.py's decoder layer). */
int flatq_add_rmsnorm(const void* h_f16, const void* y_add_f16, const void* weight_f16, int64_t rows, int64_t K, float mean_dim,
                      float eps, void* h_out_f16, void* normed_f16, flatq_stream_t stream);
int flatq_rope(void* x_f16, const void* cos_f16, const void* sin_f16, int64_t tokens, int heads, int head_dim,
               flatq_stream_t stream);
int flatq_silu_mul(const void* gu_f16, int64_t rows, int64_t inter, void* out_f16, flatq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FLATQ_B200_H_ */
