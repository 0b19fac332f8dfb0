/*
 * b200_decode.h — C ABI of libb200decode.so
 *
 * The drop-in boundary of the B200-native decode hot path.
 * Every entry point takes raw DEVICE pointers, int64 shapes/strides counted in
 * ELEMENTS (not bytes), plain scalars and the CUDA stream to launch on.  No
 * torch types cross this boundary.  Every function returns B200_OK (0) or a
 * negative b200_status; nothing here aborts the process, allocates device
 * memory behind the caller's back, synchronises the host with the device or
 * reads device metadata on the host, so every call is CUDA-graph capturable
 * (the reference contract: src/engine/model_runner.cpp:141-210).
 *
 * Each declaration cites the reference interface it replaces (paths relative
 * to the ScaleLLM tree @ffee4ffd).  INTEGRATION.md shows the reference-side
 * binding (C++ shim with the reference's own signatures, and the ctypes stub).
 */
#ifndef B200_DECODE_H_
#define B200_DECODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200_API
#else
#define B200_API __attribute__((visibility("default")))
#endif

/* Opaque cudaStream_t (CUstream).  0 / NULL = legacy default stream. */
typedef void* b200_stream_t;

typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_INVALID_ARG = -1,  /* bad shape / alignment / null pointer          */
  B200_ERR_UNSUPPORTED = -2,  /* valid in the reference, not implemented here  */
  B200_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed             */
  B200_ERR_WORKSPACE = -4     /* caller's workspace too small                  */
} b200_status;

/* Element types of activations / caches. */
typedef enum b200_dtype {
  B200_BF16 = 0,
  B200_FP16 = 1,
  B200_FP32 = 2 /* elementwise + norm kernels only */
} b200_dtype;

/* Library ABI version (bumped on any signature change). */
B200_API int b200_abi_version(void);

/* Human readable description of the last error on the calling thread. */
B200_API const char* b200_last_error(void);

/* Number of kernels this library launched on the calling thread since the
 * last reset (bench.py's "gpu_launches" claim is read from here). */
B200_API int64_t b200_launch_count(void);
B200_API void b200_launch_count_reset(void);

/* ------------------------------------------------------------------------ *
 * A4  RMSNorm                       src/kernels/layernorm_kernels.h:6-19
 *     out = T(x * rsqrt(mean(x^2) + eps)) * w          (layernorm_kernels.cu:15-41)
 *     residual form: r = r + x (stored), then the same on r   (:125-155)
 *     in/out/residual: [rows, n] contiguous; weight: [n].
 * ------------------------------------------------------------------------ */
B200_API int b200_rms_norm(void* out, const void* in, const void* weight,
                           int64_t rows, int64_t n, float eps, int dtype,
                           b200_stream_t stream);

B200_API int b200_rms_norm_residual(void* out, void* residual, const void* in,
                                    const void* weight, int64_t rows, int64_t n,
                                    float eps, int dtype, b200_stream_t stream);

/* ------------------------------------------------------------------------ *
 * A3  Rotary embedding, in place    src/kernels/pos_embedding_kernels.h:7-13
 *     q: [n_tokens, n_heads, head_dim]   (token stride q_stride, heads dense)
 *     k: [n_tokens, n_kv_heads, head_dim](token stride k_stride, heads dense)
 *     positions: [n_tokens] int32
 *     cos_sin:  [max_pos, rotary_dim] = [cos(rotary_dim/2) | sin(rotary_dim/2)]
 *     every multiply / add rounded to the element type (pos_embedding_kernels.cu:24-29)
 * ------------------------------------------------------------------------ */
B200_API int b200_rope_inplace(void* q, void* k, const int32_t* positions,
                               const void* cos_sin, int64_t n_tokens,
                               int64_t n_heads, int64_t n_kv_heads,
                               int64_t head_dim, int64_t rotary_dim,
                               int64_t q_stride, int64_t k_stride,
                               int interleaved, int dtype, b200_stream_t stream);

/* Debug / test hook (host arithmetic only): the work partition b200_paged_attn_decode would use.
 * out[8] = {impl (0 simt, 1 mma, 2 stream), n_splits, tiles per warp, padded tiles per
 * (sequence, row block, kv head), number of those, row blocks, total tiles, block-table window}. */
B200_API int b200_debug_attn_plan(int64_t batch, int max_q_len, int max_kv_len, int n_heads,
                                  int n_kv_heads, int head_dim, int block_size, int64_t* out);

/* ------------------------------------------------------------------------ *
 * A2  KV-cache slot write           src/kernels/kv_cache_kernels.h:6-11
 *     cache[slot_ids[t], h, :] = {k,v}[t, h, :]        (bit exact)
 *     k,v: [n_tokens, n_kv_heads, head_dim], token strides k_stride/v_stride
 *     caches: [n_slots, n_kv_heads, head_dim] contiguous
 * ------------------------------------------------------------------------ */
B200_API int b200_kv_write(const int32_t* slot_ids, const void* k, const void* v,
                           void* k_cache, void* v_cache, int64_t n_tokens,
                           int64_t n_kv_heads, int64_t head_dim,
                           int64_t k_stride, int64_t v_stride, int dtype,
                           b200_stream_t stream);

/* Fusion of A3 + A2 (ScaleAttnHandler::apply_pos_emb + append_kv_cache,
 * src/layers/attention/scale_attn_handler.cpp:31-42,71-84): rotates q and k in
 * place AND scatters rotated k plus v into the cache in one pass.  Results are
 * bit-identical to b200_rope_inplace followed by b200_kv_write. */
B200_API int b200_rope_kv_write(void* q, void* k, const void* v,
                                const int32_t* positions, const void* cos_sin,
                                const int32_t* slot_ids, void* k_cache,
                                void* v_cache, int64_t n_tokens, int64_t n_heads,
                                int64_t n_kv_heads, int64_t head_dim,
                                int64_t rotary_dim, int64_t q_stride,
                                int64_t k_stride, int64_t v_stride,
                                int interleaved, int dtype, b200_stream_t stream);

/* Gather (test/debug helper; KVCache::get_kv_cache, src/memory/kv_cache.cpp:60-98):
 * {k,v}_out[t] = cache[slot_ids[t]]. */
B200_API int b200_kv_gather(const int32_t* slot_ids, const void* k_cache,
                            const void* v_cache, void* k_out, void* v_out,
                            int64_t n_tokens, int64_t n_kv_heads,
                            int64_t head_dim, int dtype, b200_stream_t stream);

/* ------------------------------------------------------------------------ *
 * A5  Activations                   src/kernels/activation_kernels.h:6-14
 *     silu:      out[r, i] = T(x / (1 + exp(-x)))            in: [rows, n], row stride in_stride
 *     silu_mul:  out[r, i] = T(silu(in[r, i])) * in[r, n+i]   in: [rows, 2n] contiguous
 *     (two roundings, activation_kernels.cu:44-50,84-95)
 * ------------------------------------------------------------------------ */
B200_API int b200_silu(void* out, const void* in, int64_t rows, int64_t n,
                       int64_t in_stride, int dtype, b200_stream_t stream);

B200_API int b200_silu_mul(void* out, const void* in, int64_t rows, int64_t n,
                           int dtype, b200_stream_t stream);

/* silu(gate) * up with gate and up given as separate strided views — the exact
 * op pair Llama's MLP issues (src/models/meta/llama.h:61-64): kernel::silu then
 * a torch multiply.  gate/up: [rows, n] with row strides; out contiguous. */
B200_API int b200_silu_mul_strided(void* out, const void* gate, const void* up,
                                   int64_t rows, int64_t n, int64_t gate_stride,
                                   int64_t up_stride, int dtype,
                                   b200_stream_t stream);
/* Adjacent operators of the same reference link target (:kernels), so that a link-time swap also
 * resolves for the reference's Gemma / GPT-2 / Phi models — bit-identical to the reference kernels,
 * not on the benchmarked path:
 *   gemma_rms_norm  out = (T)(x * rstd * (1.0 + w))           src/kernels/layernorm_kernels.cu:66-123
 *   layer_norm      out = (T)((x - mean) * rstd * w (+ b))     src/kernels/layernorm_kernels.cu:185-260
 *   gelu            act 1 = gelu_new, 2 = gelu_fast (tanh.approx), optionally * up
 *                                                             src/kernels/activation_kernels.cu:13-41,108-145 */
B200_API int b200_gemma_rms_norm(void* out, const void* in, const void* weight, int64_t rows,
                                 int64_t n, float eps, int dtype, b200_stream_t stream);
B200_API int b200_layer_norm(void* out, const void* in, const void* weight, const void* bias /*nullable*/,
                             int64_t rows, int64_t n, float eps, int dtype, b200_stream_t stream);
B200_API int b200_gelu(void* out, const void* in, int64_t rows, int64_t n, int64_t in_stride, int act,
                       int with_mul, int dtype, b200_stream_t stream);


/* ------------------------------------------------------------------------ *
 * A1  Paged-KV variable-length attention (decode-shaped)
 *                                   src/kernels/attention/attn_api.h:12-27
 *     O = softmax(mask(softcap(Q K^T * sm_scale) + alibi)) V per sequence,
 *     causal with diagonal kv_len - q_len, slot lookup
 *     block_table[block_cu_lens[b] + (idx >> log2(bs))] + (idx & (bs-1))
 *     (block_table holds FIRST-SLOT ids: block_id * block_size;
 *      src/kernels/attention/kernel/sm80_kernel_mha.cuh:148-152, src/engine/batch.cpp:206-209)
 *
 *     out, q:  [n_tokens, n_heads, head_dim], strides (q_stride_t, q_stride_h, 1)
 *     caches:  [n_slots, n_kv_heads, head_dim], strides (kv_stride_s, kv_stride_h, 1)
 *     head_dim in {64, 128, 256} (tensor-core kernels) or {32, 96} (CUDA-core kernel; the
 *     reference pads these into its 64 / 128 tiles, common/static_dispatch.h:16-46); bf16 / fp16
 *     q_cu_lens, kv_cu_lens, block_cu_lens: [batch+1] int32;  block_table int32
 *     alibi_slopes: [n_heads] float32 or NULL
 *     sliding_window < 0 disables the local mask; logits_soft_cap == 0 disables it.
 *     workspace: device scratch for split-KV partials, at least
 *     b200_paged_attn_workspace_bytes(...) bytes, 16-byte aligned (may be NULL
 *     when that function returns 0).
 *     Every query token of a sequence is processed as its own work item, which is
 *     the right shape for decode / speculative decode (max_q_len <= ~8); large
 *     q_len (prefill) is correct but not the tuned path.
 * ------------------------------------------------------------------------ */
B200_API int64_t b200_paged_attn_workspace_bytes(int64_t batch, int64_t max_q_len,
                                                 int64_t max_kv_len,
                                                 int64_t n_heads,
                                                 int64_t n_kv_heads,
                                                 int64_t head_dim);

B200_API int b200_paged_attn_decode(
    void* out, const void* q, const void* k_cache, const void* v_cache,
    const int32_t* q_cu_lens, const int32_t* kv_cu_lens,
    const int32_t* block_table, const int32_t* block_cu_lens,
    const float* alibi_slopes, int64_t batch, int64_t n_heads,
    int64_t n_kv_heads, int64_t head_dim, int64_t n_slots, int64_t q_stride_t,
    int64_t q_stride_h, int64_t o_stride_t, int64_t o_stride_h,
    int64_t kv_stride_s, int64_t kv_stride_h, int block_size, int max_q_len,
    int max_kv_len, float sm_scale, float logits_soft_cap, int sliding_window,
    void* workspace, int64_t workspace_bytes, int dtype, b200_stream_t stream);

/* ------------------------------------------------------------------------ *
 * A6/A7  int4 weight x bf16 activation matmul (W4A16)
 *     replaces marlin::awq_repack / gptq_repack + marlin::gptq_gemm
 *     (src/kernels/quantization/marlin.h:17-37) behind the qlinear plugins
 *     (src/layers/quantization/qlinear_awq_marlin_impl.cpp:99-125,332-365).
 *
 *     C[M,N] = A[M,K] * W,   W[k,n] = bf16_mul( bf16(q[k,n]) - bf16(z[k/g,n]), s[k/g,n] )
 *     (exact subtract, one bf16 rounding in the multiply: marlin/numeric_conversion.h:144-167,221-240),
 *     fp32 accumulate, fp32 cross-CTA reduction, one final rounding to bf16.
 *
 *     Checkpoint formats accepted by the prepack calls (device pointers):
 *       AWQ : qweight [K, N/8] int32, nibbles along N in order [0,2,4,6,1,3,5,7];
 *             qzeros [K/g, N/8] int32 same packing; scales [K/g, N] bf16
 *             (tests/kernels/quant_utils.py:177-197)
 *       GPTQ: qweight [K/8, N] int32, nibbles along K in natural order
 *             (quant_utils.py:101-115); symmetric zero point 8 when qzeros == NULL
 *             (qlinear_gptq_marlin_impl.cpp:18-20), else qzeros [K/g, N/8] int32
 *             natural order with the GPTQ-v1 "+1" convention applied iff
 *             zeros_plus_one != 0 (src/layers/quantization/qlinear_impl.cpp:44,86);
 *             g_idx (act-order) is not supported: B200_ERR_UNSUPPORTED.
 *     group_size g in {32, 64, 128, -1(=K)};  K % 128 == 0, N % 128 == 0.
 *
 *     The prepacked buffer is an opaque tile-blob layout private to this
 *     library (DESIGN.md "W4 tile blob"); size from b200_w4a16_packed_bytes.
 * ------------------------------------------------------------------------ */
B200_API int64_t b200_w4a16_packed_bytes(int64_t K, int64_t N, int group_size);

B200_API int b200_w4a16_prepack_awq(void* packed, const int32_t* qweight,
                                    const int32_t* qzeros, const void* scales,
                                    int64_t K, int64_t N, int group_size,
                                    b200_stream_t stream);

B200_API int b200_w4a16_prepack_gptq(void* packed, const int32_t* qweight,
                                     const int32_t* qzeros, const void* scales,
                                     int64_t K, int64_t N, int group_size,
                                     int zeros_plus_one, b200_stream_t stream);

/* GPTQ act-order (desc_act): the rows of the packed weight are the checkpoint's rows sorted by
 * quant group — perm = argsort(g_idx), g_idx_sorted = g_idx[perm] (qlinear_gptq_marlin_impl.cpp:43-56,
 * marlin/gptq_repack.cu:17); the caller feeds the GEMM activations with the same column order
 * (b200_permute_cols, the reference's permute_cols_kernel marlin/gptq_gemm.cu:66-104).  Whole K
 * only (is_k_full): every group then has exactly group_size rows. */
B200_API int b200_w4a16_prepack_gptq_actorder(void* packed, const int32_t* qweight,
                                              const int32_t* qzeros, const void* scales,
                                              const int32_t* perm, const int32_t* g_idx_sorted,
                                              int64_t K, int64_t N, int group_size,
                                              int zeros_plus_one, b200_stream_t stream);
/* The operator-level drop-in for marlin::awq_repack(q_weight, out, num_bits) /
 * gptq_repack(q_weight, perm, out, num_bits) (src/kernels/quantization/marlin.h:30-37): `out` has
 * the byte count of q_weight (K * N / 2) and receives the nibble part of every tile blob, tile
 * (nt, kt) at byte (nt * K/128 + kt) * 8192; perm = act-order row order or NULL. */
B200_API int b200_w4a16_repack_awq(void* out, const int32_t* qweight, int64_t K, int64_t N,
                                   b200_stream_t stream);
B200_API int b200_w4a16_repack_gptq(void* out, const int32_t* qweight, const int32_t* perm,
                                    int64_t K, int64_t N, b200_stream_t stream);
/* ... and for what marlin::gptq_gemm receives beside it (marlin.h:17-28): scales [K/g, N] and
 * zero points [K/g, N/8] ALREADY in Marlin's column order (the layer permuted them,
 * qlinear_awq_marlin_impl.cpp:62-124); zeros_marlin NULL = symmetric, zero point 8 (has_zp
 * false).  Writes the full tile blobs b200_w4a16_gemm streams (b200_w4a16_packed_bytes). */
B200_API int b200_w4a16_assemble_marlin(void* packed, const void* nibbles, const void* scales_marlin,
                                        const int32_t* zeros_marlin, int64_t K, int64_t N,
                                        int group_size, b200_stream_t stream);
/* out[r, j] = in[r, perm[j]], bf16 / fp16 rows (strides in elements): the activation side of
 * act-order, permute_cols_kernel marlin/gptq_gemm.cu:66-104. */
B200_API int b200_permute_cols(void* out, const void* in, const int32_t* perm, int64_t rows,
                               int64_t cols, int64_t in_stride, int64_t out_stride, int dtype,
                               b200_stream_t stream);
/* Inverse of the prepack (debug / parity): W_out[K,N] bf16 = dequantised weights. */
B200_API int b200_w4a16_dequant(void* w_out, const void* packed, int64_t K,
                                int64_t N, int group_size, b200_stream_t stream);

/* Workspace of b200_w4a16_gemm: the fp32 stream-K partials of one chunk of <= 128 rows
 * (slots x min(M,128) x N floats).  No initialisation contract: it is fully overwritten before
 * it is read (unlike Marlin's zeroed lock workspace, marlin.h:24, which has no counterpart here). */
B200_API int64_t b200_w4a16_workspace_bytes(int64_t M, int64_t N, int64_t K);

/* A: [M, K] bf16 row stride lda;  C: [M, N] bf16 row stride ldc; bias: [N] bf16 or NULL.
 * Two launches: the stream-K GEMM (fp32 partials into the workspace) and a reduction pass. */
B200_API int b200_w4a16_gemm(void* C, const void* A, const void* packed,
                             const void* bias, int64_t M, int64_t N, int64_t K,
                             int64_t lda, int64_t ldc, int group_size,
                             void* workspace, int64_t workspace_bytes,
                             b200_stream_t stream);

/* "Partials" mode (B200-native fusion of the GEMM's cross-CTA reduction into its consumer).
 * The GEMM is stream-K: the (n tile, k tile) units are cut into equal contiguous shares, one per
 * CTA, and every CTA writes the fp32 partial of each tile it touches to
 * partials[slot][M][N], slot = its rank among that tile's contributors.  The consumer
 * (b200_rms_norm_residual_splitk, b200_ar_allreduce_splitk, or the reduction pass of
 * b200_w4a16_gemm) recomputes the same partition from (N, K), sums each tile's slots in slot
 * order and rounds once to the element type — exactly what a GEMM epilogue would have stored.
 * Replaces the o_proj / down_proj + residual add + RMSNorm sequence of
 * models/meta/llama.h:170-177.  b200_w4a16_splitk_splits returns the slot count the partials
 * buffer must have for a [K, N] weight on the current device (<= 8).  M <= 128. */
B200_API int b200_w4a16_splitk_splits(int64_t M, int64_t N, int64_t K);
B200_API int b200_w4a16_gemm_splitk(float* partials, const void* A, const void* packed, int64_t M,
                                    int64_t N, int64_t K, int64_t lda, int group_size, int splits,
                                    b200_stream_t stream);
/* C[M, N] bf16 (row stride ldc) = bf16(sum_slots partials) (+ bias): the plain reduction pass. */
B200_API int b200_w4a16_reduce_partials(void* C, const float* partials, int splits, int64_t gemm_k,
                                        const void* bias, int64_t M, int64_t N, int64_t ldc,
                                        b200_stream_t stream);
/* residual += T(sum_slots partials); out = rms_norm(residual) * weight.
 * partials: [splits, rows, n] fp32 from a GEMM with reduction dimension gemm_k. */
B200_API int b200_rms_norm_residual_splitk(void* out, void* residual, const float* partials,
                                           int splits, int64_t gemm_k, const void* weight,
                                           int64_t rows, int64_t n, float eps, int dtype,
                                           b200_stream_t stream);

/* Fused consumers of the qkv / gate_up GEMM partials (same partition rules as above):
 *   b200_rope_kv_write_splitk: qkv[T, (H + 2 Hkv) D] = T(sum_slots partials), then exactly
 *     b200_rope_kv_write on its q | k | v column blocks (q, k rotated in place, rotated k and v
 *     scattered to their cache slots).  Replaces qkv_proj's epilogue + apply_rotary_pos_emb +
 *     set_kv_cache (models/meta/llama.h:123-133).
 *   b200_silu_mul_splitk: out[rows, inter] = silu(T(gate)) * T(up), gate | up the two halves of the
 *     gate_up GEMM row (models/meta/llama.h:61-64, activation.cpp:101-103 rounding). */
B200_API int b200_rope_kv_write_splitk(void* qkv, const float* partials, int splits, int64_t gemm_k,
                                       const int32_t* positions, const void* cos_sin,
                                       const int32_t* slot_ids, void* k_cache, void* v_cache,
                                       int64_t n_tokens, int64_t n_heads, int64_t n_kv_heads,
                                       int64_t head_dim, int64_t rotary_dim, int interleaved,
                                       int dtype, b200_stream_t stream);
B200_API int b200_silu_mul_splitk(void* out, const float* partials, int splits, int64_t gemm_k,
                                  int64_t rows, int64_t inter, int dtype, b200_stream_t stream);

/* Debug / test hook (host arithmetic only, no device): the stream-K partition the GEMM and its
 * consumers use for a [K, N] weight split over `ctas` CTAs with `nsub` (1 or 2) weight tiles per
 * partition tile.  plan_out[5] = {units, ctas used, k tiles, partition tiles, slots};
 * first_owner_out / contrib_out (optional, one entry per partition tile) = first contributing CTA
 * and number of contributors of that tile (the slot a CTA writes is its index minus first_owner). */
B200_API int b200_debug_w4a16_plan(int64_t N, int64_t K, int ctas, int nsub, int32_t* plan_out,
                                   int32_t* first_owner_out, int32_t* contrib_out);

/* ------------------------------------------------------------------------ *
 * A8  Dense bf16 linear, decode-sized batches
 *     replaces F::linear -> cuBLASLt of ColumnParallelLinearImpl / RowParallelLinearImpl::forward
 *     (src/layers/linear/parallel_linear.cpp:256-263,294-308) and lm_head
 *     (src/models/meta/llama.h:259-265): C[M, N] = A[M, K] W[N, K]^T (+ bias), bf16 in / out,
 *     fp32 accumulation, one rounding.  W is the nn.Linear weight as the checkpoint stores it
 *     ([N, K] row-major, row stride ldw elements); N %% 8 == 0, K %% 64 == 0; any M (one pass over
 *     W per 128 rows).  workspace: b200_dense_workspace_bytes (fp32 stream-K partials).
 * ------------------------------------------------------------------------ */
B200_API int64_t b200_dense_workspace_bytes(int64_t M, int64_t N, int64_t K);
B200_API int b200_dense_gemm(void* C, const void* A, const void* W, const void* bias /*nullable [N]*/,
                             int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw, int64_t ldc,
                             void* workspace, int64_t workspace_bytes, b200_stream_t stream);
/* Partials mode, same contract as b200_w4a16_gemm_splitk (the reduction, and whatever follows it in
 * the layer — RoPE + KV write, residual + RMSNorm, SiLU*mul, the TP all-reduce — runs in the
 * consumer): partials [b200_dense_splitk_splits()][M][N] fp32, M <= 128, N and K multiples of 128;
 * the stream-K partition is the int4 GEMM's for the same (N, K), so every b200_*_splitk consumer
 * takes these partials unchanged. */
B200_API int b200_dense_splitk_splits(int64_t M, int64_t N, int64_t K);
B200_API int b200_dense_gemm_splitk(float* partials, const void* A, const void* W, int64_t M, int64_t N,
                                    int64_t K, int64_t lda, int64_t ldw, int splits, b200_stream_t stream);

/* Debug hook: when non-NULL, every b200_w4a16_gemm CTA records clock64() milestones into
 * device_buffer[blockIdx.x * 16 + slot] (long long).  Pass NULL to disable (default). */
B200_API void b200_debug_set_trace(void* device_buffer);

/* ------------------------------------------------------------------------ *
 * Greedy sampling tail (the step driver's last kernel):
 *     out[r] = argmax_j logits[r, j]   (first index of the maximum, NaN counts as the maximum:
 *     torch.argmax semantics, which the reference's greedy path uses, src/sampling/sampler.cpp).
 * ------------------------------------------------------------------------ */
B200_API int b200_argmax(int64_t* out, const void* logits, int64_t rows, int64_t n, int64_t stride,
                         int dtype, b200_stream_t stream);

/* The logits processors of the sampling tail, in place on logits [batch, vocab] (contiguous), with
 * the semantics and rounding points of src/kernels/sampling/penalty_kernels.cu:9-33,52-75,107-140
 * and softmax_kernels.cu:11-54 (declared in src/kernels/sampling/sampling_kernels.h:7-29):
 *   temperature   logits[b, :] *= (t[b] == 0 ? 1 : 1 / t[b])
 *   repetition    for the lens[b] unique ids of row b: x < 0 ? x * p[b] : x / p[b]
 *   freq/presence for ids with count > 0: x -= count * freq[b]; x -= presence[b]
 *   softmax       exp(x - max) stored in T, summed from the stored values, / (sum + 1e-6)
 * token_ids int64 [batch, max_len], token_counts / token_ids_lens int32. */
B200_API int b200_apply_temperature(void* logits, const void* temperatures, int64_t batch,
                                    int64_t vocab, int dtype, b200_stream_t stream);
B200_API int b200_apply_repetition_penalty(void* logits, const int64_t* token_ids,
                                           const int32_t* token_ids_lens, const void* penalties,
                                           int64_t batch, int64_t vocab, int64_t max_len, int dtype,
                                           b200_stream_t stream);
B200_API int b200_apply_frequency_presence_penalty(void* logits, const int64_t* token_ids,
                                                   const int32_t* token_counts,
                                                   const int32_t* token_ids_lens,
                                                   const void* frequency_penalties,
                                                   const void* presence_penalties, int64_t batch,
                                                   int64_t vocab, int64_t max_len, int dtype,
                                                   b200_stream_t stream);
B200_API int b200_softmax(void* logits, int64_t batch, int64_t vocab, int dtype, b200_stream_t stream);
/* top-k / top-p filter, in place: replaces TopKTopPLogitsProcessor::forward
 * (src/sampling/logits_processor.h:243-276: sort descending, mask sorted positions >= top_k[b], softmax of
 * the rest, mask positions whose exclusive cumulative probability exceeds top_p[b], scatter back — a
 * [batch, vocab] sort and four more library launches) by one launch that finds the cut with a radix
 * histogram and writes -inf over everything outside it.  top_k [batch] int64 (<= 0: no top-k limit),
 * top_p [batch] float (>= 1: no top-p limit); either may be NULL.  logits [batch, vocab] bf16 / fp16, row
 * stride in elements.  Of equal logits straddling the cut the lowest vocabulary indices are kept. */
B200_API int b200_topk_topp_filter(void* logits, const int64_t* top_k /*nullable*/, const float* top_p /*nullable*/,
                                   int64_t batch, int64_t vocab, int64_t stride, int dtype, b200_stream_t stream);

/* ------------------------------------------------------------------------ *
 * A9  Tensor-parallel all-reduce over NVLink peer memory
 *     replaces ProcessGroupNCCL::allreduce (src/model_parallel/process_group.cpp:135-153)
 *     for the <= 1 MiB row-parallel reductions of the decode step (two-shot row-partitioned
 *     push protocol above two ranks, one-shot pull at two; csrc/allreduce.cu).
 *
 *     One communicator per GPU.  One process per GPU: every rank calls
 *     b200_ar_create (allocates its symmetric buffer + flags and returns an IPC
 *     handle blob), exchanges the blobs out of band (torch.distributed
 *     all_gather in this repo), then calls b200_ar_open_peers with all ranks'
 *     blobs in rank order.  All ranks in one process, one thread per GPU (the
 *     reference engine): b200_ar_create_all.
 * ------------------------------------------------------------------------ */
typedef struct b200_ar_comm b200_ar_comm;
#define B200_AR_HANDLE_BYTES 128

B200_API int b200_ar_create(b200_ar_comm** comm, int rank, int world_size,
                            int64_t max_bytes, void* handle_out /*[B200_AR_HANDLE_BYTES]*/);
B200_API int b200_ar_open_peers(b200_ar_comm* comm,
                                const void* all_handles /*[world][B200_AR_HANDLE_BYTES]*/);
/* All ranks in ONE process, one thread per GPU — the reference engine's model:
 * ProcessGroup::create_process_groups -> ncclCommInitAll (process_group.cpp:98-118).  Creates
 * comms[r] on devices[r] for r = 0..world_size-1 and maps every peer's region through CUDA peer
 * access (no IPC, no handle exchange); all or nothing.  Each communicator is then used from the
 * thread that has its device current, exactly like the IPC form. */
B200_API int b200_ar_create_all(b200_ar_comm** comms /*[world_size] out*/, const int* devices,
                                int world_size, int64_t max_bytes);
/* In-place sum of data[count] (bf16/fp16/fp32) over all ranks, on `stream`. */
B200_API int b200_ar_allreduce(b200_ar_comm* comm, void* data, int64_t count,
                               int dtype, b200_stream_t stream);
/* All-gather along the last dimension: in [rows, row_bytes] per rank -> out [rows, world *
 * row_bytes] with rank r's row at byte offset r * row_bytes — what
 * gather_from_model_parallel_region builds with allgather + cat(dim=-1)
 * (src/model_parallel/model_parallel.cpp:13-31).  Bit exact; rows * row_bytes <= max_bytes,
 * row_bytes % 16 == 0; shares the communicator's epoch with the all-reduces (every rank must
 * issue the same sequence of collectives). */
B200_API int b200_ar_allgather(b200_ar_comm* comm, void* out, const void* in, int64_t rows,
                               int64_t row_bytes, b200_stream_t stream);
/* Same reduction, but this rank's input is the producing GEMM's stream-K partials
 * [splits][count] fp32 (b200_w4a16_gemm_splitk of a [gemm_k, n] weight, count = rows * n): the
 * copy-in stage sums each tile's slots and rounds once, so the row-parallel GEMM needs no
 * reduction pass of its own.  out: [count] bf16/fp16. */
B200_API int b200_ar_allreduce_splitk(b200_ar_comm* comm, void* out, const float* partials,
                                      int splits, int64_t gemm_k, int64_t n, int64_t count,
                                      int dtype, b200_stream_t stream);
/* ... and with the consumer fused as well: residual += T(all-reduced row); out = rms_norm(residual)
 * * weight, one launch for the row-parallel GEMM's reduction, the TP all-reduce, the residual add
 * and the RMSNorm (models/meta/llama.h:170-177 under tensor parallelism).  Two-shot form (default
 * above two ranks): rows <= 128, n <= 8192; one-shot form: rows <= 64, n <= 4096. */
B200_API int b200_ar_allreduce_splitk_norm(b200_ar_comm* comm, void* out, void* residual,
                                           const float* partials, int splits, int64_t gemm_k,
                                           const void* weight, int64_t rows, int64_t n, float eps,
                                           int dtype, b200_stream_t stream);
/* Greedy sampling over a column-parallel (vocabulary-sharded) lm_head without gathering the
 * logits: out[r] = argmax over all ranks' columns of row r, exactly what
 * torch.argmax(gather_from_model_parallel_region(logits), -1) gives (lm_head gather_output=true,
 * src/models/meta/llama.h:259-265 -> model_parallel.cpp:13-31, then the greedy sampler): first
 * index of the maximum, NaN counts as the maximum.  logits: this rank's [rows, n_local] shard
 * (row stride `stride` elements), global column = rank * n_local + local column.  One launch: an
 * 8-byte candidate per rank and row crosses NVLink.  rows <= 128.  Shares the communicator's
 * epoch with the other collectives. */
B200_API int b200_ar_argmax(b200_ar_comm* comm, int64_t* out, const void* logits, int64_t rows,
                            int64_t n_local, int64_t stride, int dtype, b200_stream_t stream);
B200_API int b200_ar_destroy(b200_ar_comm* comm);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* B200_DECODE_H_ */
