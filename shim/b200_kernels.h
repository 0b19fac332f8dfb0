// b200_kernels.h — C++ drop-in for ScaleLLM's operator-level kernel API.
//
// Same namespaces, names, signatures, borrowing and in-place conventions as the reference headers
// these replace (link this translation unit instead of the reference's :kernels,
// :attention.kernels and :marlin.kernels targets; tests/test_shim.py compiles a translation unit
// against the reference's OWN headers and links it with this library):
//
//   llm::kernel::rms_norm / rms_norm_residual      src/kernels/layernorm_kernels.h:6-19
//   llm::kernel::apply_rotary_pos_emb              src/kernels/pos_embedding_kernels.h:7-13
//   llm::kernel::set_kv_cache                      src/kernels/kv_cache_kernels.h:6-11
//   llm::kernel::silu / silu_with_mul              src/kernels/activation_kernels.h:6-14
//   llm::paged_kv_varlen_mha                       src/kernels/attention/attn_api.h:12-27
//   marlin::awq_repack / gptq_repack / gptq_gemm   src/kernels/quantization/marlin.h:17-37
//
// Every function forwards to the C ABI of libb200decode.so (include/b200_decode.h) on
// at::cuda::getCurrentCUDAStream().  Errors follow the reference: a failed call is fatal
// (TORCH_CHECK -> c10::Error, where the reference LOG(FATAL)s).
#pragma once

#include <torch/torch.h>

#include <optional>

namespace llm::kernel {

void rms_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, float epsilon);

void rms_norm_residual(torch::Tensor& out, torch::Tensor& residual, torch::Tensor input,
                       torch::Tensor weight, float epsilon);

void apply_rotary_pos_emb(torch::Tensor& querys, torch::Tensor& keys,
                          const torch::Tensor& positions, const torch::Tensor& cos_sin,
                          int rotary_dim, bool interleaved);

void set_kv_cache(const torch::Tensor& slot_ids, const torch::Tensor& keys,
                  const torch::Tensor& values, torch::Tensor& key_cache,
                  torch::Tensor& value_cache);

torch::Tensor silu(torch::Tensor input);
torch::Tensor silu_with_mul(torch::Tensor input);

// the rest of the reference's :kernels surface (adjacent to the Llama path: Gemma, GPT-2, Phi):
// src/kernels/layernorm_kernels.h:11-25, src/kernels/activation_kernels.h:8-13
void gemma_rms_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, float epsilon);
void layer_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, torch::Tensor bias,
                float epsilon);
torch::Tensor gelu_new(torch::Tensor input);
torch::Tensor gelu_fast(torch::Tensor input);
torch::Tensor gelu_new_with_mul(torch::Tensor input);
torch::Tensor gelu_fast_with_mul(torch::Tensor input);

// the sampling tail's logits processors, in place (src/kernels/sampling/sampling_kernels.h:7-29;
// the top-k / top-p sampler of that header is not replaced: greedy selection is b200_argmax)
void apply_temperature_penalty(torch::Tensor& logits, const torch::Tensor& temperatures);
void apply_repetition_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                              const torch::Tensor& token_ids_lens, const torch::Tensor& penalities);
void apply_frequency_presence_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                                      const torch::Tensor& token_counts,
                                      const torch::Tensor& token_ids_lens,
                                      const torch::Tensor& frequency_penalties,
                                      const torch::Tensor& presence_penalties);
void invoke_softmax(torch::Tensor& logits);
// B200 extension: TopKTopPLogitsProcessor::forward (src/sampling/logits_processor.h:243-276) in place in
// one launch instead of sort + masked_fill + softmax + cumsum + gather.  top_k: int64 [n] (<= 0: no limit)
// or undefined; top_p: float [n] (>= 1: no limit) or undefined.  The processor's forward() becomes
//   auto out = logits.clone(); kernel::apply_top_k_top_p(out, top_k, top_p); return out;
void apply_top_k_top_p(torch::Tensor& logits, const torch::Tensor& top_k, const torch::Tensor& top_p);

// B200 extension used by B200AttnHandler: rope + cache write in one launch
// (bit-identical to apply_rotary_pos_emb followed by set_kv_cache).
void rope_and_set_kv_cache(torch::Tensor& querys, torch::Tensor& keys, const torch::Tensor& values,
                           const torch::Tensor& positions, const torch::Tensor& cos_sin,
                           const torch::Tensor& slot_ids, torch::Tensor& key_cache,
                           torch::Tensor& value_cache, int rotary_dim, bool interleaved);

}  // namespace llm::kernel

namespace llm {

void paged_kv_varlen_mha(torch::Tensor& out, const torch::Tensor& query,
                         const torch::Tensor& key_cache, const torch::Tensor& value_cache,
                         const torch::Tensor& q_cu_lens, const torch::Tensor& kv_cu_lens,
                         const torch::Tensor& block_table, const torch::Tensor& block_cu_lens,
                         const std::optional<torch::Tensor>& alibi_slopes, int block_size,
                         int max_q_len, int max_kv_len, float sm_scale, float logits_soft_cap,
                         int sliding_window);

}  // namespace llm

namespace marlin {

// ---- the reference's EXACT signatures (src/kernels/quantization/marlin.h:17-37) ----------------
// A ScaleLLM build that links this library instead of :marlin.kernels needs no source edit in
// qlinear_awq_marlin_impl.cpp / qlinear_gptq_marlin_impl.cpp:
//   * awq_repack / gptq_repack fill `out` — (K/16, N*16/8) int32, the byte count of q_weight — with
//     the nibble part of this library's tile blobs (opaque to the caller, who only hands it back);
//     gptq_repack's `perm` (act-order row order, may be empty) is honoured;
//   * gptq_gemm takes the scales [K/g, N] and zero points [K/g, N/8] the layer permuted into
//     Marlin's column order, assembles the full tile blobs once per weight (cached by the identity
//     of B / scales / zeros) and runs b200_w4a16_gemm; `zeros` is read iff has_zp (else the
//     symmetric zero point 8), g_idx / perm select act-order (activation columns are gathered by
//     `perm` first; whole-K only: is_k_full must be true), `workspace` is left zeroed as Marlin
//     promises, use_fp32_reduce is what this library always does.  num_bits must be 4; A, C and
//     scales must be bf16 (the reference also takes fp16: refused here, not reinterpreted).
void awq_repack(const torch::Tensor& q_weight, torch::Tensor& out, int64_t num_bits);

void gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& perm, torch::Tensor& out,
                 int64_t num_bits);

void gptq_gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C,
               const torch::Tensor& scales, const torch::Tensor& zeros, const torch::Tensor& g_idx,
               const torch::Tensor& perm, torch::Tensor& workspace, int num_bits, bool is_k_full,
               bool has_zp, bool use_fp32_reduce);

// ---- B200 extensions (not in the reference): one-step repack into the full tile blobs ----------
// `out` must hold b200_packed_bytes(K, N, group_size) bytes; checkpoint-order scales / zero points
// travel inside the blob.  gptq_gemm recognises such a B by its size and then ignores its
// scales / zeros arguments.
void b200_awq_repack(const torch::Tensor& q_weight, const torch::Tensor& q_zeros,
                     const torch::Tensor& scales, torch::Tensor& out, int64_t group_size);

void b200_gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& scales,
                      torch::Tensor& out, int64_t group_size);

// bytes of the full tile blobs / of the workspace b200_w4a16_gemm needs
int64_t b200_packed_bytes(int64_t K, int64_t N, int64_t group_size);
int64_t b200_workspace_bytes(int64_t M, int64_t N, int64_t K);
// blobs assembled by gptq_gemm so far (tests)
int64_t b200_assembled_weights();

}  // namespace marlin
