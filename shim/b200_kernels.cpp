// b200_kernels.cpp — see b200_kernels.h.  Thin torch::Tensor -> C-ABI forwarding; no compute here.
#include "b200_kernels.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "../include/b200_decode.h"

namespace {

int dtype_of(const torch::Tensor& t) {
  switch (t.scalar_type()) {
    case torch::kBFloat16: return B200_BF16;
    case torch::kHalf: return B200_FP16;
    case torch::kFloat: return B200_FP32;
    default: TORCH_CHECK(false, "b200: unsupported dtype ", t.scalar_type());
  }
}

b200_stream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void ok(int rc, const char* what) {
  TORCH_CHECK(rc == B200_OK, "b200 ", what, " failed (", rc, "): ", b200_last_error());
}

const int32_t* i32(const torch::Tensor& t) {
  TORCH_CHECK(t.scalar_type() == torch::kInt && t.is_contiguous(), "expected contiguous int32");
  return t.const_data_ptr<int32_t>();
}

}  // namespace

namespace llm::kernel {

void rms_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, float epsilon) {
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous(), "tensors must be contiguous");
  const int64_t n = input.size(-1);
  if (input.numel() == 0) return;
  ok(b200_rms_norm(out.data_ptr(), input.const_data_ptr(), weight.const_data_ptr(),
                   input.numel() / n, n, epsilon, dtype_of(input), stream()),
     "rms_norm");
}

void rms_norm_residual(torch::Tensor& out, torch::Tensor& residual, torch::Tensor input,
                       torch::Tensor weight, float epsilon) {
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous() && residual.is_contiguous());
  const int64_t n = input.size(-1);
  if (input.numel() == 0) return;
  ok(b200_rms_norm_residual(out.data_ptr(), residual.data_ptr(), input.const_data_ptr(),
                            weight.const_data_ptr(), input.numel() / n, n, epsilon,
                            dtype_of(input), stream()),
     "rms_norm_residual");
}

void apply_rotary_pos_emb(torch::Tensor& querys, torch::Tensor& keys,
                          const torch::Tensor& positions, const torch::Tensor& cos_sin,
                          int rotary_dim, bool interleaved) {
  TORCH_CHECK(querys.stride(-1) == 1 && querys.stride(-2) == querys.size(-1));
  TORCH_CHECK(keys.stride(-1) == 1 && keys.stride(-2) == keys.size(-1));
  ok(b200_rope_inplace(querys.data_ptr(), keys.data_ptr(), i32(positions),
                       cos_sin.const_data_ptr(), querys.size(-3), querys.size(-2), keys.size(-2),
                       querys.size(-1), rotary_dim, querys.stride(-3), keys.stride(-3),
                       interleaved ? 1 : 0, dtype_of(querys), stream()),
     "apply_rotary_pos_emb");
}

void set_kv_cache(const torch::Tensor& slot_ids, const torch::Tensor& keys,
                  const torch::Tensor& values, torch::Tensor& key_cache,
                  torch::Tensor& value_cache) {
  TORCH_CHECK(keys.stride(-1) == 1 && keys.stride(-2) == keys.size(-1));
  TORCH_CHECK(values.stride(-1) == 1 && values.stride(-2) == values.size(-1));
  ok(b200_kv_write(i32(slot_ids), keys.const_data_ptr(), values.const_data_ptr(),
                   key_cache.data_ptr(), value_cache.data_ptr(), keys.size(-3), keys.size(-2),
                   keys.size(-1), keys.stride(-3), values.stride(-3), dtype_of(keys), stream()),
     "set_kv_cache");
}

void rope_and_set_kv_cache(torch::Tensor& querys, torch::Tensor& keys, const torch::Tensor& values,
                           const torch::Tensor& positions, const torch::Tensor& cos_sin,
                           const torch::Tensor& slot_ids, torch::Tensor& key_cache,
                           torch::Tensor& value_cache, int rotary_dim, bool interleaved) {
  ok(b200_rope_kv_write(querys.data_ptr(), keys.data_ptr(), values.const_data_ptr(),
                        i32(positions), cos_sin.const_data_ptr(), i32(slot_ids),
                        key_cache.data_ptr(), value_cache.data_ptr(), querys.size(-3),
                        querys.size(-2), keys.size(-2), querys.size(-1), rotary_dim,
                        querys.stride(-3), keys.stride(-3), values.stride(-3), interleaved ? 1 : 0,
                        dtype_of(querys), stream()),
     "rope_and_set_kv_cache");
}

torch::Tensor silu(torch::Tensor input) {
  TORCH_CHECK(input.dim() == 2 && input.stride(1) == 1);
  torch::Tensor out = torch::empty({input.size(0), input.size(1)}, input.options());
  ok(b200_silu(out.data_ptr(), input.const_data_ptr(), input.size(0), input.size(1),
               input.stride(0), dtype_of(input), stream()),
     "silu");
  return out;
}

torch::Tensor silu_with_mul(torch::Tensor input) {
  TORCH_CHECK(input.is_contiguous() && input.dim() == 2);
  const int64_t n = input.size(1) / 2;
  torch::Tensor out = torch::empty({input.size(0), n}, input.options());
  ok(b200_silu_mul(out.data_ptr(), input.const_data_ptr(), input.size(0), n, dtype_of(input),
                   stream()),
     "silu_with_mul");
  return out;
}

}  // namespace llm::kernel

namespace llm {

void paged_kv_varlen_mha(torch::Tensor& out, const torch::Tensor& query,
                         const torch::Tensor& key_cache, const torch::Tensor& value_cache,
                         const torch::Tensor& q_cu_lens, const torch::Tensor& kv_cu_lens,
                         const torch::Tensor& block_table, const torch::Tensor& block_cu_lens,
                         const std::optional<torch::Tensor>& alibi_slopes, int block_size,
                         int max_q_len, int max_kv_len, float sm_scale, float logits_soft_cap,
                         int sliding_window) {
  const int64_t batch = q_cu_lens.size(0) - 1;
  const int64_t n_heads = query.size(-2), head_dim = query.size(-1);
  const int64_t n_kv_heads = key_cache.size(-2);
  const int64_t ws_bytes =
      b200_paged_attn_workspace_bytes(batch, max_q_len, max_kv_len, n_heads, n_kv_heads, head_dim);
  // split-KV scratch comes from the torch caching allocator (graph-capture safe, handler.h:19-23
  // lets a handler ask for a workspace; the operator-level call allocates like Marlin's c_tmp)
  torch::Tensor ws;
  if (ws_bytes > 0) ws = torch::empty({ws_bytes}, query.options().dtype(torch::kByte));
  ok(b200_paged_attn_decode(
         out.data_ptr(), query.const_data_ptr(), key_cache.const_data_ptr(),
         value_cache.const_data_ptr(), i32(q_cu_lens), i32(kv_cu_lens), i32(block_table),
         i32(block_cu_lens),
         alibi_slopes.has_value() ? alibi_slopes->const_data_ptr<float>() : nullptr, batch,
         n_heads, n_kv_heads, head_dim, key_cache.size(0), query.stride(0), query.stride(1),
         out.stride(0), out.stride(1), key_cache.stride(0), key_cache.stride(1), block_size,
         max_q_len, max_kv_len, sm_scale, logits_soft_cap, sliding_window,
         ws_bytes > 0 ? ws.data_ptr() : nullptr, ws_bytes, dtype_of(query), stream()),
     "paged_kv_varlen_mha");
}

}  // namespace llm

namespace marlin {

int64_t b200_packed_bytes(int64_t K, int64_t N, int64_t group_size) {
  return b200_w4a16_packed_bytes(K, N, static_cast<int>(group_size));
}
int64_t b200_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return b200_w4a16_workspace_bytes(M, N, K);
}

void awq_repack(const torch::Tensor& q_weight, const torch::Tensor& q_zeros,
                const torch::Tensor& scales, torch::Tensor& out, int64_t group_size) {
  const int64_t K = q_weight.size(0), N = q_weight.size(1) * 8;
  TORCH_CHECK(out.numel() * out.element_size() >= b200_packed_bytes(K, N, group_size),
              "awq_repack: out too small for the B200 tile-blob layout");
  ok(b200_w4a16_prepack_awq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(),
                            q_zeros.const_data_ptr<int32_t>(), scales.const_data_ptr(), K, N,
                            static_cast<int>(group_size), stream()),
     "awq_repack");
}

void gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& scales, torch::Tensor& out,
                 int64_t group_size) {
  const int64_t K = q_weight.size(0) * 8, N = q_weight.size(1);
  TORCH_CHECK(out.numel() * out.element_size() >= b200_packed_bytes(K, N, group_size),
              "gptq_repack: out too small for the B200 tile-blob layout");
  ok(b200_w4a16_prepack_gptq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(), nullptr,
                             scales.const_data_ptr(), K, N, static_cast<int>(group_size), 0,
                             stream()),
     "gptq_repack");
}

void gptq_gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C,
               const torch::Tensor& scales, const torch::Tensor& /*zeros*/,
               const torch::Tensor& g_idx, const torch::Tensor& /*perm*/, torch::Tensor& workspace,
               int num_bits, bool /*is_k_full*/, bool /*has_zp*/, bool /*use_fp32_reduce*/) {
  TORCH_CHECK(num_bits == 4, "b200 gptq_gemm: 4-bit weights only");
  TORCH_CHECK(g_idx.numel() == 0, "b200 gptq_gemm: act-order (g_idx) is not supported");
  const int64_t M = A.size(0), K = A.size(1), N = C.size(1);
  const int64_t groups = scales.size(0);
  const int group_size = groups <= 1 ? -1 : static_cast<int>(K / groups);
  ok(b200_w4a16_gemm(C.data_ptr(), A.const_data_ptr(), B.const_data_ptr(), nullptr, M, N, K,
                     A.stride(0), C.stride(0), group_size, workspace.data_ptr(),
                     workspace.numel() * workspace.element_size(), stream()),
     "gptq_gemm");
}

}  // namespace marlin
