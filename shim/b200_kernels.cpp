// b200_kernels.cpp — see b200_kernels.h.  Thin torch::Tensor -> C-ABI forwarding; no compute here.
#include "b200_kernels.h"

#include <cuda_runtime_api.h>

#include <mutex>
#include <vector>

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

void gemma_rms_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, float epsilon) {
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous(), "gemma_rms_norm: contiguous tensors");
  const int64_t n = input.size(-1);
  ok(b200_gemma_rms_norm(out.data_ptr(), input.const_data_ptr(), weight.const_data_ptr(),
                         input.numel() / n, n, epsilon, dtype_of(input), stream()),
     "gemma_rms_norm");
}

void layer_norm(torch::Tensor& out, torch::Tensor input, torch::Tensor weight, torch::Tensor bias,
                float epsilon) {
  TORCH_CHECK(input.is_contiguous() && out.is_contiguous(), "layer_norm: contiguous tensors");
  const int64_t n = input.size(-1);
  ok(b200_layer_norm(out.data_ptr(), input.const_data_ptr(), weight.const_data_ptr(),
                     bias.defined() ? bias.const_data_ptr() : nullptr, input.numel() / n, n, epsilon,
                     dtype_of(input), stream()),
     "layer_norm");
}

namespace {
torch::Tensor gelu_impl(const torch::Tensor& input, int act, bool with_mul) {
  TORCH_CHECK(input.dim() == 2 && input.stride(1) == 1 && (!with_mul || input.is_contiguous()),
              "gelu: [rows, n] with dense rows (contiguous for the fused multiply)");
  const int64_t n = with_mul ? input.size(1) / 2 : input.size(1);
  torch::Tensor out = torch::empty({input.size(0), n}, input.options());
  ok(b200_gelu(out.data_ptr(), input.const_data_ptr(), input.size(0), n, input.stride(0), act,
               with_mul ? 1 : 0, dtype_of(input), stream()),
     "gelu");
  return out;
}
}  // namespace
void apply_temperature_penalty(torch::Tensor& logits, const torch::Tensor& temperatures) {
  TORCH_CHECK(logits.dim() == 2 && logits.is_contiguous() && temperatures.is_contiguous() &&
                  temperatures.scalar_type() == logits.scalar_type(),
              "apply_temperature_penalty: contiguous [batch, vocab] logits, temperatures of the same dtype");
  ok(b200_apply_temperature(logits.data_ptr(), temperatures.const_data_ptr(), logits.size(0), logits.size(1),
                            dtype_of(logits), stream()),
     "apply_temperature_penalty");
}

void apply_repetition_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                              const torch::Tensor& token_ids_lens, const torch::Tensor& penalities) {
  TORCH_CHECK(logits.dim() == 2 && logits.is_contiguous() && token_ids.is_contiguous() &&
                  token_ids.scalar_type() == torch::kLong && token_ids_lens.scalar_type() == torch::kInt,
              "apply_repetition_penalty: int64 token_ids [batch, max_len], int32 lens");
  ok(b200_apply_repetition_penalty(logits.data_ptr(), token_ids.const_data_ptr<int64_t>(),
                                   token_ids_lens.const_data_ptr<int32_t>(), penalities.const_data_ptr(),
                                   logits.size(0), logits.size(1), token_ids.size(1), dtype_of(logits), stream()),
     "apply_repetition_penalty");
}

void apply_frequency_presence_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                                      const torch::Tensor& token_counts,
                                      const torch::Tensor& token_ids_lens,
                                      const torch::Tensor& frequency_penalties,
                                      const torch::Tensor& presence_penalties) {
  TORCH_CHECK(logits.dim() == 2 && logits.is_contiguous() && token_ids.is_contiguous() &&
                  token_counts.is_contiguous() && token_ids.scalar_type() == torch::kLong &&
                  token_counts.scalar_type() == torch::kInt && token_ids_lens.scalar_type() == torch::kInt,
              "apply_frequency_presence_penalty: int64 ids, int32 counts / lens");
  ok(b200_apply_frequency_presence_penalty(
         logits.data_ptr(), token_ids.const_data_ptr<int64_t>(), token_counts.const_data_ptr<int32_t>(),
         token_ids_lens.const_data_ptr<int32_t>(), frequency_penalties.const_data_ptr(),
         presence_penalties.const_data_ptr(), logits.size(0), logits.size(1), token_ids.size(1),
         dtype_of(logits), stream()),
     "apply_frequency_presence_penalty");
}

void invoke_softmax(torch::Tensor& logits) {
  TORCH_CHECK(logits.dim() == 2 && logits.is_contiguous(), "invoke_softmax: contiguous [batch, vocab]");
  ok(b200_softmax(logits.data_ptr(), logits.size(0), logits.size(1), dtype_of(logits), stream()), "softmax");
}

void apply_top_k_top_p(torch::Tensor& logits, const torch::Tensor& top_k, const torch::Tensor& top_p) {
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1, "apply_top_k_top_p: [batch, vocab] logits, dense rows");
  TORCH_CHECK(logits.scalar_type() == torch::kBFloat16 || logits.scalar_type() == torch::kHalf,
              "apply_top_k_top_p: bf16 / fp16 logits");
  torch::Tensor k, p;
  if (top_k.defined()) {
    k = top_k.to(logits.device(), torch::kInt64).reshape({-1}).contiguous();
    TORCH_CHECK(k.numel() == logits.size(0), "apply_top_k_top_p: one top_k per row");
  }
  if (top_p.defined()) {
    p = top_p.to(logits.device(), torch::kFloat32).reshape({-1}).contiguous();
    TORCH_CHECK(p.numel() == logits.size(0), "apply_top_k_top_p: one top_p per row");
  }
  ok(b200_topk_topp_filter(logits.data_ptr(), k.defined() ? k.const_data_ptr<int64_t>() : nullptr,
                           p.defined() ? p.const_data_ptr<float>() : nullptr, logits.size(0), logits.size(1),
                           logits.stride(0), dtype_of(logits), stream()),
     "topk_topp_filter");
}

torch::Tensor gelu_new(torch::Tensor input) { return gelu_impl(input, 1, false); }
torch::Tensor gelu_fast(torch::Tensor input) { return gelu_impl(input, 2, false); }
torch::Tensor gelu_new_with_mul(torch::Tensor input) { return gelu_impl(input, 1, true); }
torch::Tensor gelu_fast_with_mul(torch::Tensor input) { return gelu_impl(input, 2, true); }

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

void b200_awq_repack(const torch::Tensor& q_weight, const torch::Tensor& q_zeros,
                     const torch::Tensor& scales, torch::Tensor& out, int64_t group_size) {
  const int64_t K = q_weight.size(0), N = q_weight.size(1) * 8;
  TORCH_CHECK(scales.scalar_type() == torch::kBFloat16, "awq_repack: bf16 scales only, got ", scales.scalar_type());
  TORCH_CHECK(out.numel() * out.element_size() >= b200_packed_bytes(K, N, group_size),
              "awq_repack: out too small for the B200 tile-blob layout");
  ok(b200_w4a16_prepack_awq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(),
                            q_zeros.const_data_ptr<int32_t>(), scales.const_data_ptr(), K, N,
                            static_cast<int>(group_size), stream()),
     "awq_repack");
}

void b200_gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& scales,
                      torch::Tensor& out, int64_t group_size) {
  const int64_t K = q_weight.size(0) * 8, N = q_weight.size(1);
  TORCH_CHECK(scales.scalar_type() == torch::kBFloat16, "gptq_repack: bf16 scales only, got ", scales.scalar_type());
  TORCH_CHECK(out.numel() * out.element_size() >= b200_packed_bytes(K, N, group_size),
              "gptq_repack: out too small for the B200 tile-blob layout");
  ok(b200_w4a16_prepack_gptq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(), nullptr,
                             scales.const_data_ptr(), K, N, static_cast<int>(group_size), 0,
                             stream()),
     "gptq_repack");
}

// ---- the reference's exact signatures --------------------------------------------------------
void awq_repack(const torch::Tensor& q_weight, torch::Tensor& out, int64_t num_bits) {
  TORCH_CHECK(num_bits == 4, "b200 awq_repack: 4-bit weights only");
  TORCH_CHECK(q_weight.is_cuda() && q_weight.is_contiguous() && q_weight.scalar_type() == torch::kInt,
              "awq_repack: q_weight must be a contiguous int32 CUDA tensor");
  const int64_t K = q_weight.size(0), N = q_weight.size(1) * 8;
  TORCH_CHECK(out.is_contiguous() && out.numel() * out.element_size() == K * N / 2,
              "awq_repack: out must be (K/16, N*16/8) int32");
  ok(b200_w4a16_repack_awq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(), K, N, stream()),
     "awq_repack");
}

void gptq_repack(const torch::Tensor& q_weight, const torch::Tensor& perm, torch::Tensor& out,
                 int64_t num_bits) {
  TORCH_CHECK(num_bits == 4, "b200 gptq_repack: 4-bit weights only");
  TORCH_CHECK(q_weight.is_cuda() && q_weight.is_contiguous() && q_weight.scalar_type() == torch::kInt,
              "gptq_repack: q_weight must be a contiguous int32 CUDA tensor");
  const int64_t K = q_weight.size(0) * 8, N = q_weight.size(1);
  TORCH_CHECK(out.is_contiguous() && out.numel() * out.element_size() == K * N / 2,
              "gptq_repack: out must be (K/16, N*16/8) int32");
  const int32_t* pp = nullptr;
  if (perm.defined() && perm.numel() > 0) {
    TORCH_CHECK(perm.numel() == K && perm.scalar_type() == torch::kInt && perm.is_contiguous(),
                "gptq_repack: perm must be a contiguous int32 tensor of K entries");
    pp = perm.const_data_ptr<int32_t>();
  }
  ok(b200_w4a16_repack_gptq(out.data_ptr(), q_weight.const_data_ptr<int32_t>(), pp, K, N, stream()),
     "gptq_repack");
}

namespace {
// Full tile blobs assembled from (nibble tiles, Marlin-order scales, Marlin-packed zero points),
// once per weight.  An entry is tied to the identity of the three tensors: it holds weak references
// to their TensorImpls (so the addresses cannot be reused while the entry lives) and is dropped
// when any of them has died.
struct Assembled {
  c10::weak_intrusive_ptr<c10::TensorImpl> b, s, z;
  const void *b_ptr, *s_ptr, *z_ptr;
  torch::Tensor packed;
};
std::mutex g_asm_mu;
std::vector<Assembled> g_asm;

torch::Tensor assembled_for(const torch::Tensor& B, const torch::Tensor& scales,
                            const torch::Tensor& zeros, bool has_zp, int64_t K, int64_t N,
                            int group_size) {
  const void* zp = has_zp ? zeros.const_data_ptr() : nullptr;
  {
    std::lock_guard<std::mutex> lk(g_asm_mu);
    for (size_t i = 0; i < g_asm.size();) {
      auto& e = g_asm[i];
      if (e.b.expired() || e.s.expired() || (e.z_ptr && e.z.expired())) {
        g_asm.erase(g_asm.begin() + static_cast<long>(i));
        continue;
      }
      if (e.b_ptr == B.const_data_ptr() && e.s_ptr == scales.const_data_ptr() && e.z_ptr == zp)
        return e.packed;
      ++i;
    }
  }
  auto packed = torch::empty({b200_packed_bytes(K, N, group_size)}, B.options().dtype(torch::kByte));
  ok(b200_w4a16_assemble_marlin(packed.data_ptr(), B.const_data_ptr(), scales.const_data_ptr(),
                                static_cast<const int32_t*>(zp), K, N, group_size, stream()),
     "assemble_marlin");
  // the GEMM's weight producer does not wait on the predecessor kernel (weights are constants)
  TORCH_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream())) == cudaSuccess,
              "gptq_gemm: stream synchronize after the one-time weight assembly failed");
  std::lock_guard<std::mutex> lk(g_asm_mu);
  g_asm.push_back({c10::weak_intrusive_ptr<c10::TensorImpl>(B.getIntrusivePtr()),
                   c10::weak_intrusive_ptr<c10::TensorImpl>(scales.getIntrusivePtr()),
                   has_zp ? c10::weak_intrusive_ptr<c10::TensorImpl>(zeros.getIntrusivePtr())
                          : c10::weak_intrusive_ptr<c10::TensorImpl>(B.getIntrusivePtr()),
                   B.const_data_ptr(), scales.const_data_ptr(), zp, packed});
  return packed;
}
}  // namespace

int64_t b200_assembled_weights() {
  std::lock_guard<std::mutex> lk(g_asm_mu);
  return static_cast<int64_t>(g_asm.size());
}

void gptq_gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C,
               const torch::Tensor& scales, const torch::Tensor& zeros,
               const torch::Tensor& g_idx, const torch::Tensor& perm, torch::Tensor& workspace,
               int num_bits, bool is_k_full, bool has_zp, bool /*use_fp32_reduce*/) {
  (void)workspace;  // Marlin's lock counters: zero on entry, untouched here, so zero on exit
  TORCH_CHECK(num_bits == 4, "b200 gptq_gemm: 4-bit weights only");
  TORCH_CHECK(A.dim() == 2 && C.dim() == 2 && A.size(0) == C.size(0), "gptq_gemm: A (m, k), C (m, n)");
  // bf16 only: the reference's Marlin also takes fp16, which these kernels would reinterpret
  TORCH_CHECK(A.scalar_type() == torch::kBFloat16 && C.scalar_type() == torch::kBFloat16,
              "b200 gptq_gemm: bf16 activations / output only, got ", A.scalar_type(), " / ", C.scalar_type());
  TORCH_CHECK(A.stride(1) == 1 && C.stride(1) == 1, "gptq_gemm: A and C rows must be dense");
  const int64_t M = A.size(0), K = A.size(1), N = C.size(1);
  const int64_t b_bytes = B.numel() * static_cast<int64_t>(B.element_size());
  torch::Tensor a = A;
  const bool act_order = g_idx.defined() && g_idx.numel() > 0;
  if (act_order) {
    TORCH_CHECK(is_k_full, "b200 gptq_gemm: act-order on a K-sharded weight (is_k_full = false) is not supported");
    TORCH_CHECK(perm.defined() && perm.numel() == K && perm.scalar_type() == torch::kInt,
                "gptq_gemm: act-order needs perm (k) int32");
    a = torch::empty({M, K}, A.options());
    ok(b200_permute_cols(a.data_ptr(), A.const_data_ptr(), perm.const_data_ptr<int32_t>(), M, K,
                         A.stride(0), a.stride(0), B200_BF16, stream()),
       "permute_cols");
  }
  const void* packed = nullptr;
  torch::Tensor held;
  int group_size = -1;
  if (b_bytes == K * N / 2) {  // the reference's layout contract: nibble tiles + Marlin-order scales / zeros
    TORCH_CHECK(scales.scalar_type() == torch::kBFloat16, "b200 gptq_gemm: bf16 scales only, got ", scales.scalar_type());
    TORCH_CHECK(scales.dim() == 2 && scales.size(1) == N, "gptq_gemm: scales (n_groups, n)");
    const int64_t groups = scales.size(0);
    group_size = groups <= 1 ? -1 : static_cast<int>(K / groups);
    if (has_zp)
      TORCH_CHECK(zeros.defined() && zeros.numel() == groups * N / 8 && zeros.scalar_type() == torch::kInt,
                  "gptq_gemm: has_zp needs zeros (n_groups, n/8) int32");
    held = assembled_for(B, scales, zeros, has_zp, K, N, group_size);
    packed = held.const_data_ptr();
  } else {                      // B200 extension: B already holds the full tile blobs
    const int64_t groups = scales.defined() && scales.dim() == 2 ? scales.size(0) : 1;
    group_size = groups <= 1 ? -1 : static_cast<int>(K / groups);
    TORCH_CHECK(b_bytes >= b200_packed_bytes(K, N, group_size), "gptq_gemm: B is neither (K/16, N*16/8) int32 nor a tile-blob buffer");
    packed = B.const_data_ptr();
  }
  if (M == 0) return;
  // partials workspace from the caching allocator (graph-capture safe, never shared between calls)
  auto ws = torch::empty({b200_workspace_bytes(M, N, K)}, A.options().dtype(torch::kByte));
  ok(b200_w4a16_gemm(C.data_ptr(), a.const_data_ptr(), packed, nullptr, M, N, K, a.stride(0),
                     C.stride(0), group_size, ws.data_ptr(), ws.numel(), stream()),
     "gptq_gemm");
}

}  // namespace marlin
