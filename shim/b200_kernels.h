// b200_kernels.h — C++ drop-in for ScaleLLM's operator-level kernel API.
//
// Same namespaces, names, argument order, borrowing and in-place conventions as the reference
// headers these replace (link this translation unit instead of the reference's :kernels,
// :attention.kernels and :marlin.kernels targets):
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

// The repack output is this library's tile-blob layout, not Marlin's fragment layout; it is
// opaque to the callers (qlinear_*_marlin_impl.cpp only hands it back to gptq_gemm), but it is
// (K/128)*(N/128)*(8192+groups) bytes — callers must size `out` with b200_w4a16_packed_bytes
// (the scales / zero points travel inside the blob, so the permuted scale and zero-point
// tensors of the Marlin path are no longer read).
void awq_repack(const torch::Tensor& q_weight, const torch::Tensor& q_zeros,
                const torch::Tensor& scales, torch::Tensor& out, int64_t group_size);

void gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& scales, torch::Tensor& out,
                 int64_t group_size);

void gptq_gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C,
               const torch::Tensor& scales, const torch::Tensor& zeros, const torch::Tensor& g_idx,
               const torch::Tensor& perm, torch::Tensor& workspace, int num_bits, bool is_k_full,
               bool has_zp, bool use_fp32_reduce);

// bytes of the packed weight / of the workspace gptq_gemm needs (no initialisation needed)
int64_t b200_packed_bytes(int64_t K, int64_t N, int64_t group_size);
int64_t b200_workspace_bytes(int64_t M, int64_t N, int64_t K);

}  // namespace marlin
