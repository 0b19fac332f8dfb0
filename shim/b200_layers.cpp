// b200_layers.cpp — see b200_layers.h.  Host logic only: shapes, ownership, call order; every
// device operation is a C-ABI call into libb200decode (or a torch library op for the dense
// embedding lookup / lm_head, which this path does not replace).
#include "b200_layers.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdlib>

#include "../include/b200_decode.h"
#include "b200_kernels.h"

namespace {

int dtype_of(torch::ScalarType t) {
  switch (t) {
    case torch::kBFloat16: return B200_BF16;
    case torch::kHalf: return B200_FP16;
    case torch::kFloat: return B200_FP32;
    default: TORCH_CHECK(false, "b200: unsupported dtype ", t);
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
const torch::Tensor& need(const llm::StateDict& sd, const std::string& key) {
  auto it = sd.find(key);
  TORCH_CHECK(it != sd.end(), "state dict has no entry '", key, "'");
  return it->second;
}

}  // namespace

namespace llm {

// ---------------------------------------------------------------------------------------------
// fused operator-level entry points
// ---------------------------------------------------------------------------------------------
namespace kernel {

W4Partials w4a16_gemm_partials(const torch::Tensor& a, const torch::Tensor& packed, int64_t N,
                               int64_t group_size) {
  TORCH_CHECK(a.dim() == 2 && a.scalar_type() == torch::kBFloat16 && a.stride(1) == 1,
              "w4a16_gemm_partials: A must be [M, K] bf16 with unit inner stride");
  const int64_t M = a.size(0), K = a.size(1);
  const int slots = b200_w4a16_splitk_splits(M, N, K);
  TORCH_CHECK(slots >= 1, "w4a16_gemm_partials: bad shape K=", K, " N=", N);
  W4Partials p{torch::empty({slots, M, N}, a.options().dtype(torch::kFloat)), K};
  ok(b200_w4a16_gemm_splitk(p.data.data_ptr<float>(), a.const_data_ptr(), packed.const_data_ptr(), M,
                            N, K, a.stride(0), static_cast<int>(group_size), slots, stream()),
     "w4a16_gemm_partials");
  return p;
}

torch::Tensor w4a16_reduce_partials(const W4Partials& p, const std::optional<torch::Tensor>& bias) {
  const int64_t S = p.data.size(0), M = p.data.size(1), N = p.data.size(2);
  torch::Tensor out = torch::empty({M, N}, p.data.options().dtype(torch::kBFloat16));
  ok(b200_w4a16_reduce_partials(out.data_ptr(), p.data.const_data_ptr<float>(), static_cast<int>(S),
                                p.K, bias.has_value() ? bias->const_data_ptr() : nullptr, M, N, N,
                                stream()),
     "w4a16_reduce_partials");
  return out;
}

torch::Tensor rms_norm_residual_partials(const W4Partials& p, torch::Tensor& residual,
                                         const torch::Tensor& weight, float eps) {
  TORCH_CHECK(residual.is_contiguous() && p.data.is_contiguous());
  const int64_t S = p.data.size(0), rows = p.data.size(1), n = p.data.size(2);
  torch::Tensor out = torch::empty_like(residual);
  if (rows == 0) return out;
  ok(b200_rms_norm_residual_splitk(out.data_ptr(), residual.data_ptr(),
                                   p.data.const_data_ptr<float>(), static_cast<int>(S), p.K,
                                   weight.const_data_ptr(), rows, n, eps,
                                   dtype_of(residual.scalar_type()), stream()),
     "rms_norm_residual_partials");
  return out;
}

torch::Tensor rope_and_set_kv_cache_partials(const W4Partials& p, int64_t n_heads,
                                             int64_t n_kv_heads, int64_t head_dim,
                                             const torch::Tensor& positions,
                                             const torch::Tensor& cos_sin,
                                             const torch::Tensor& slot_ids, torch::Tensor& key_cache,
                                             torch::Tensor& value_cache, int rotary_dim,
                                             bool interleaved, torch::ScalarType dtype) {
  const int64_t S = p.data.size(0), T = p.data.size(1), n = p.data.size(2);
  TORCH_CHECK(n == (n_heads + 2 * n_kv_heads) * head_dim, "qkv partials have the wrong row length");
  torch::Tensor qkv = torch::empty({T, n}, p.data.options().dtype(dtype));
  ok(b200_rope_kv_write_splitk(qkv.data_ptr(), p.data.const_data_ptr<float>(), static_cast<int>(S),
                               p.K, i32(positions), cos_sin.const_data_ptr(), i32(slot_ids),
                               key_cache.data_ptr(), value_cache.data_ptr(), T, n_heads, n_kv_heads,
                               head_dim, rotary_dim, interleaved ? 1 : 0, dtype_of(dtype), stream()),
     "rope_and_set_kv_cache_partials");
  return qkv;
}

torch::Tensor silu_mul_partials(const W4Partials& p, torch::ScalarType dtype) {
  const int64_t S = p.data.size(0), rows = p.data.size(1), n2 = p.data.size(2);
  torch::Tensor out = torch::empty({rows, n2 / 2}, p.data.options().dtype(dtype));
  ok(b200_silu_mul_splitk(out.data_ptr(), p.data.const_data_ptr<float>(), static_cast<int>(S), p.K,
                          rows, n2 / 2, dtype_of(dtype), stream()),
     "silu_mul_partials");
  return out;
}

torch::Tensor argmax(const torch::Tensor& logits) {
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1);
  torch::Tensor out = torch::empty({logits.size(0)}, logits.options().dtype(torch::kLong));
  ok(b200_argmax(out.data_ptr<int64_t>(), logits.const_data_ptr(), logits.size(0), logits.size(1),
                 logits.stride(0), dtype_of(logits.scalar_type()), stream()),
     "argmax");
  return out;
}

// dense bf16 linear, F::linear(x, W[N, K]) of the reference's dense layers and lm_head
// (parallel_linear.cpp:258,299; llama.h:259-265): csrc/dense.cu
torch::Tensor dense_linear(const torch::Tensor& x, const torch::Tensor& w) {
  TORCH_CHECK(x.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16 && w.dim() == 2 &&
                  x.size(-1) == w.size(1) && w.stride(1) == 1,
              "dense_linear: bf16 x [.., K] and W [N, K]");
  torch::Tensor x2 = x.reshape({-1, x.size(-1)});
  if (x2.stride(1) != 1 || x2.stride(0) % 8 != 0) x2 = x2.contiguous();
  const int64_t M = x2.size(0), K = x2.size(1), N = w.size(0);
  torch::Tensor out = torch::empty({M, N}, x.options());
  if (M == 0) return out;
  torch::Tensor ws = torch::empty({b200_dense_workspace_bytes(M, N, K)}, x.options().dtype(torch::kByte));
  ok(b200_dense_gemm(out.data_ptr(), x2.const_data_ptr(), w.const_data_ptr(), nullptr, M, N, K, x2.stride(0),
                     w.stride(0), out.stride(0), ws.data_ptr(), ws.numel(), stream()),
     "dense_gemm");
  auto shape = x.sizes().vec();
  shape.back() = N;
  return out.view(shape);
}

}  // namespace kernel

// ---------------------------------------------------------------------------------------------
// KV cache
// ---------------------------------------------------------------------------------------------
void KVCache::set_kv_cache(const torch::Tensor& slot_ids, const torch::Tensor& keys,
                           const torch::Tensor& values) {
  kernel::set_kv_cache(slot_ids, keys, values, key_cache_, value_cache_);
}

// ---------------------------------------------------------------------------------------------
// attention handler
// ---------------------------------------------------------------------------------------------
B200Handler::B200Handler(float sm_scale, float logits_soft_cap,
                         std::optional<torch::Tensor> alibi_slopes, torch::Tensor cos_sin,
                         int64_t rotary_dim, bool interleaved)
    : sm_scale_(sm_scale), soft_cap_(logits_soft_cap), alibi_(std::move(alibi_slopes)),
      cos_sin_(std::move(cos_sin)), rotary_dim_(rotary_dim), interleaved_(interleaved) {}

std::tuple<torch::Tensor, torch::Tensor> B200Handler::apply_pos_emb(const torch::Tensor& query,
                                                                    const torch::Tensor& key,
                                                                    const torch::Tensor& positions) {
  torch::Tensor q = query, k = key;
  if (has_rope() && positions.defined())
    kernel::apply_rotary_pos_emb(q, k, positions, cos_sin_, static_cast<int>(rotary_dim_),
                                 interleaved_);
  return {q, k};
}

void B200Handler::append_kv_cache(KVCache& kv_cache, const torch::Tensor& key,
                                  const torch::Tensor& value, const InputParameters& input_params) {
  if (!kv_cache.empty()) kv_cache.set_kv_cache(input_params.new_cache_slots, key, value);
}

void B200Handler::apply_pos_emb_and_append(torch::Tensor& query, torch::Tensor& key,
                                           const torch::Tensor& value, const torch::Tensor& positions,
                                           KVCache& kv_cache, const InputParameters& input_params) {
  if (!has_rope() || !positions.defined() || kv_cache.empty()) {
    apply_pos_emb(query, key, positions);
    append_kv_cache(kv_cache, key, value, input_params);
    return;
  }
  auto [kc, vc] = kv_cache.get_kv_cache();
  kernel::rope_and_set_kv_cache(query, key, value, positions, cos_sin_, input_params.new_cache_slots,
                                kc, vc, static_cast<int>(rotary_dim_), interleaved_);
}

torch::Tensor B200Handler::qkv_from_partials(const W4Partials& qkv, int64_t n_heads,
                                             int64_t n_kv_heads, int64_t head_dim,
                                             const torch::Tensor& positions, KVCache& kv_cache,
                                             const InputParameters& input_params,
                                             torch::ScalarType dtype) {
  TORCH_CHECK(has_rope() && !kv_cache.empty(), "qkv_from_partials needs RoPE and a KV cache");
  auto [kc, vc] = kv_cache.get_kv_cache();
  return kernel::rope_and_set_kv_cache_partials(qkv, n_heads, n_kv_heads, head_dim, positions,
                                                cos_sin_, input_params.new_cache_slots, kc, vc,
                                                static_cast<int>(rotary_dim_), interleaved_, dtype);
}

void B200Handler::batch_decode(const torch::Tensor& query, const KVCache& kv_cache,
                               const InputParameters& input_params, int32_t sliding_window,
                               torch::Tensor& output) {
  auto [kc, vc] = kv_cache.get_kv_cache();
  paged_kv_varlen_mha(output, query, kc, vc, input_params.q_cu_seq_lens, input_params.kv_cu_seq_lens,
                      input_params.block_tables, input_params.cu_block_lens, alibi_,
                      static_cast<int>(kv_cache.block_size()), input_params.q_max_seq_len,
                      input_params.kv_max_seq_len, sm_scale_, soft_cap_, sliding_window);
}

// ---------------------------------------------------------------------------------------------
// int4 linear
// ---------------------------------------------------------------------------------------------
QLinearB200Impl::QLinearB200Impl(int64_t in_features, int64_t out_features, bool bias,
                                 const QuantArgs& quant_args, const torch::TensorOptions& options)
    : K_(in_features), N_(out_features), qa_(quant_args), options_(options), has_bias_(bias) {
  // qlinear_awq_marlin_impl.cpp:28-31,150-151
  TORCH_CHECK(qa_.bits == 4, "B200 W4A16 path supports 4-bit weights only");
  TORCH_CHECK(qa_.group_size == -1 || qa_.group_size == 32 || qa_.group_size == 64 ||
                  qa_.group_size == 128,
              "group_size ", qa_.group_size, " not in (-1, 32, 64, 128)");
  TORCH_CHECK(!qa_.desc_act, "act-order (desc_act) checkpoints are not supported");
  TORCH_CHECK(K_ % 128 == 0 && N_ % 128 == 0,
              "W4A16 needs in_features % 128 == 0 and out_features % 128 == 0 per shard");
  TORCH_CHECK(qa_.quant_method == "awq" || qa_.quant_method == "gptq", "quant_method must be awq|gptq");
}

void QLinearB200Impl::load_state_dict(const StateDict& sd) {
  qweight_ = need(sd, "qweight").to(options_.device()).contiguous();
  scales_ = need(sd, "scales").to(options_.device()).to(torch::kBFloat16).contiguous();
  auto z = sd.find("qzeros");
  qzeros_ = z == sd.end() ? torch::Tensor() : z->second.to(options_.device()).contiguous();
  if (has_bias_) bias_ = need(sd, "bias").to(options_.device()).to(torch::kBFloat16).contiguous();
  packed_ = torch::Tensor();
}

void QLinearB200Impl::verify_loaded_weights() const {
  TORCH_CHECK(packed_.defined() || (qweight_.defined() && scales_.defined()),
              "qweight / qzeros / scales not loaded");
  TORCH_CHECK(qa_.quant_method != "awq" || packed_.defined() || qzeros_.defined(),
              "awq checkpoints need qzeros");
}

void QLinearB200Impl::ensure_packed() {
  if (packed_.defined()) return;
  verify_loaded_weights();
  const int64_t bytes = b200_w4a16_packed_bytes(K_, N_, static_cast<int>(qa_.group_size));
  TORCH_CHECK(bytes > 0, "bad W4A16 shape");
  packed_ = torch::empty({bytes}, options_.dtype(torch::kByte));
  if (qa_.quant_method == "awq") {
    ok(b200_w4a16_prepack_awq(packed_.data_ptr(), qweight_.const_data_ptr<int32_t>(),
                              qzeros_.const_data_ptr<int32_t>(), scales_.const_data_ptr(), K_, N_,
                              static_cast<int>(qa_.group_size), stream()),
       "prepack_awq");
  } else {  // GPTQ: zeros stored minus one (qlinear_impl.cpp:44); symmetric -> zero point 8
    const int32_t* qz = (qa_.is_sym || !qzeros_.defined()) ? nullptr : qzeros_.const_data_ptr<int32_t>();
    ok(b200_w4a16_prepack_gptq(packed_.data_ptr(), qweight_.const_data_ptr<int32_t>(), qz,
                               scales_.const_data_ptr(), K_, N_, static_cast<int>(qa_.group_size), 1,
                               stream()),
       "prepack_gptq");
  }
  // The first GEMM follows on the same stream with the programmatic-launch attribute and its
  // weight producer does not execute griddepcontrol.wait (weights are constants): make the
  // prepack's writes visible first.  One-time, never inside a graph capture (it allocates).
  TORCH_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream())) == cudaSuccess,
              "prepack: stream synchronize failed");
  qweight_ = qzeros_ = scales_ = torch::Tensor();  // checkpoint-format copies are no longer needed
}

torch::Tensor QLinearB200Impl::forward(torch::Tensor input) {
  // the W4A16 kernels compute in bf16 only (activations, scales, output): an fp16 tensor would be
  // reinterpreted silently.  The reference's Marlin path takes fp16 too; here it is refused.
  TORCH_CHECK(input.scalar_type() == torch::kBFloat16,
              "QLinearB200Impl: bf16 activations only, got ", input.scalar_type());
  TORCH_CHECK(input.size(-1) == K_, "QLinearB200Impl: input features ", input.size(-1), " != ", K_);
  ensure_packed();
  torch::Tensor x = input.reshape({-1, input.size(-1)});
  if (x.stride(-1) != 1) x = x.contiguous();
  const int64_t M = x.size(0);
  torch::Tensor out = torch::empty({M, N_}, x.options());
  if (M == 0) return out;
  const int64_t ws = b200_w4a16_workspace_bytes(M, N_, K_);
  if (!workspace_.defined() || workspace_.numel() < ws) {
    // grow-only: a CUDA graph captured with the smaller buffer keeps replaying into it
    if (workspace_.defined()) retired_workspaces_.push_back(workspace_);
    workspace_ = torch::empty({ws}, options_.dtype(torch::kByte));
  }
  ok(b200_w4a16_gemm(out.data_ptr(), x.const_data_ptr(), packed_.const_data_ptr(),
                     bias_.defined() ? bias_.const_data_ptr() : nullptr, M, N_, K_, x.stride(0),
                     out.stride(0), static_cast<int>(qa_.group_size), workspace_.data_ptr(),
                     workspace_.numel(), stream()),
     "w4a16_gemm");
  auto shape = input.sizes().vec();
  shape.back() = N_;
  return out.view(shape);
}

W4Partials QLinearB200Impl::forward_partials(const torch::Tensor& input) {
  ensure_packed();
  return kernel::w4a16_gemm_partials(input.reshape({-1, input.size(-1)}), packed_, N_,
                                     qa_.group_size);
}

// ---------------------------------------------------------------------------------------------
// RMSNorm
// ---------------------------------------------------------------------------------------------
RMSNormImpl::RMSNormImpl(int64_t dim, float eps, const torch::TensorOptions& options)
    : weight(torch::ones({dim}, options)), eps_(eps) {}

void RMSNormImpl::load_state_dict(const StateDict& sd) {
  weight.copy_(need(sd, "weight").to(weight.options()));
}

torch::Tensor RMSNormImpl::forward(const torch::Tensor& input) {
  torch::Tensor out = torch::empty_like(input);
  kernel::rms_norm(out, input, weight, eps_);
  return out;
}

torch::Tensor RMSNormImpl::forward_residual(const torch::Tensor& input, torch::Tensor& residual) {
  torch::Tensor out = torch::empty_like(input);
  kernel::rms_norm_residual(out, residual, input, weight, eps_);
  return out;
}

torch::Tensor RMSNormImpl::forward_residual_partials(const W4Partials& input,
                                                     torch::Tensor& residual) {
  return kernel::rms_norm_residual_partials(input, residual, weight, eps_);
}

// ---------------------------------------------------------------------------------------------
// tensor-parallel shards of checkpoint-format int4 linears
//   AWQ: qweight [K, N/8] (8 columns per word), qzeros [K/g, N/8], scales [K/g, N];
//   GPTQ: qweight [K/8, N] (8 rows per word), qzeros [K/g, N/8] (optional), scales [K/g, N]
// ---------------------------------------------------------------------------------------------
StateDict shard_qlinear_columns(const StateDict& t, const QuantArgs& qa, int64_t c0, int64_t c1) {
  TORCH_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && c0 < c1, "column shard must align to 8 columns");
  const bool awq = qa.quant_method == "awq";
  StateDict out;
  const torch::Tensor qw = need(t, "qweight");
  out["qweight"] = awq ? qw.slice(1, c0 / 8, c1 / 8) : qw.slice(1, c0, c1);
  out["scales"] = need(t, "scales").slice(1, c0, c1);
  auto z = t.find("qzeros");
  if (z != t.end()) out["qzeros"] = z->second.slice(1, c0 / 8, c1 / 8);
  return out;
}

StateDict shard_qlinear_rows(const StateDict& t, const QuantArgs& qa, int64_t k0, int64_t k1) {
  const bool awq = qa.quant_method == "awq";
  const torch::Tensor qw = need(t, "qweight");
  const int64_t K = awq ? qw.size(0) : qw.size(0) * 8;
  const int64_t g = qa.group_size > 0 ? qa.group_size : K;
  TORCH_CHECK(k0 % g == 0 && k1 % g == 0 && k0 < k1 && k1 <= K,
              "row-parallel shard must align to quant groups");  // qlinear_awq_marlin_impl.cpp:287
  StateDict out;
  out["qweight"] = awq ? qw.slice(0, k0, k1) : qw.slice(0, k0 / 8, k1 / 8);
  out["scales"] = need(t, "scales").slice(0, k0 / g, k1 / g);
  auto z = t.find("qzeros");
  if (z != t.end()) out["qzeros"] = z->second.slice(0, k0 / g, k1 / g);
  return out;
}

namespace {
StateDict cat_columns(const std::vector<StateDict>& parts) {
  StateDict out;
  for (const auto& kv : parts[0]) {
    std::vector<torch::Tensor> ts;
    for (const auto& p : parts) ts.push_back(need(p, kv.first));
    out[kv.first] = torch::cat(ts, 1);
  }
  return out;
}
}  // namespace

LlamaLayerShards shard_llama_layer(const StateDict& qkv, const StateDict& o, const StateDict& gate_up,
                                   const StateDict& down, const LlamaArgs& args, const QuantArgs& qa,
                                   int rank, int world) {
  const int64_t w = world, r = rank;
  TORCH_CHECK(w >= 1 && r >= 0 && r < w, "bad rank / world_size");
  const int64_t D = args.head_dim, H = args.n_heads, Hkv = args.n_kv_heads, I = args.intermediate_size;
  TORCH_CHECK(H % w == 0 && I % w == 0 && (Hkv % w == 0 || w % Hkv == 0), "sizes must divide by world_size");
  const int64_t Hl = H / w, Hkvl = std::max<int64_t>(1, Hkv / w), Il = I / w;
  // kv heads of this rank: a contiguous share, or one replicated head when Hkv < w
  // (qkv_parallel_linear.cpp:28-70)
  const int64_t kv0 = Hkv >= w ? r * (Hkv / w) : r / (w / Hkv);
  const int64_t q0 = r * Hl * D, k_base = H * D, v_base = (H + Hkv) * D;
  LlamaLayerShards sh;
  sh.qkv = cat_columns({shard_qlinear_columns(qkv, qa, q0, q0 + Hl * D),
                        shard_qlinear_columns(qkv, qa, k_base + kv0 * D, k_base + (kv0 + Hkvl) * D),
                        shard_qlinear_columns(qkv, qa, v_base + kv0 * D, v_base + (kv0 + Hkvl) * D)});
  sh.gate_up = cat_columns({shard_qlinear_columns(gate_up, qa, r * Il, (r + 1) * Il),
                            shard_qlinear_columns(gate_up, qa, I + r * Il, I + (r + 1) * Il)});
  sh.o = shard_qlinear_rows(o, qa, r * Hl * D, (r + 1) * Hl * D);
  sh.down = shard_qlinear_rows(down, qa, r * Il, (r + 1) * Il);
  return sh;
}

// ---------------------------------------------------------------------------------------------
// the decode step
// ---------------------------------------------------------------------------------------------
LlamaDecoderStep::LlamaDecoderStep(const LlamaArgs& args, const QuantArgs& qa,
                                   const torch::Tensor& inv_freq,
                                   const torch::TensorOptions& options,
                                   const ParallelArgs& parallel_args)
    : args_(args), quant_args_(qa), options_(options), parallel_args_(parallel_args) {
  const int64_t w = parallel_args.world_size();
  TORCH_CHECK(w >= 1 && args.n_heads % w == 0 && args.intermediate_size % w == 0 &&
                  args.hidden_size % w == 0 && args.vocab_size % w == 0,
              "heads / intermediate / hidden / vocab sizes must divide by world_size ", w);
  TORCH_CHECK(args.n_kv_heads % w == 0 || w % args.n_kv_heads == 0,
              "n_kv_heads ", args.n_kv_heads, " and world_size ", w, " must divide one another");
  if (w > 1) {
    pg_ = dynamic_cast<ProcessGroupB200*>(parallel_args.process_group());
    TORCH_CHECK(pg_ != nullptr, "tensor parallelism needs a ProcessGroupB200");
  }
  n_heads_ = args.n_heads / w;                                   // llama.h:83-90
  n_kv_heads_ = std::max<int64_t>(1, args.n_kv_heads / w);
  inter_ = args.intermediate_size / w;
  const int64_t h = args.hidden_size, D = args.head_dim;
  const int64_t q_size = n_heads_ * D, kv_size = n_kv_heads_ * D;
  const char* env = std::getenv("B200_FUSE_SPLITK");
  fuse_partials = !(env && env[0] == '0');
  // cos | sin cache in the model dtype (pos_embedding.cpp:183-215)
  torch::Tensor t = torch::arange(args.max_position_embeddings, torch::kFloat);
  torch::Tensor freqs = torch::outer(t, inv_freq.to(torch::kFloat).cpu());
  torch::Tensor cos_sin = torch::cat({freqs.cos(), freqs.sin()}, -1).to(options);
  handler_ = std::make_unique<B200Handler>(1.0f / std::sqrt(static_cast<float>(D)), 0.0f,
                                           std::nullopt, cos_sin, D, /*interleaved=*/false);
  layers_.resize(args.n_layers);
  for (auto& L : layers_) {
    L.input_norm = std::make_unique<RMSNormImpl>(h, args.rms_norm_eps, options);
    L.post_norm = std::make_unique<RMSNormImpl>(h, args.rms_norm_eps, options);
    L.qkv = std::make_unique<QLinearB200Impl>(h, q_size + 2 * kv_size, false, qa, options);
    L.o = std::make_unique<QLinearB200Impl>(q_size, h, false, qa, options);
    L.gate_up = std::make_unique<QLinearB200Impl>(h, 2 * inter_, false, qa, options);
    L.down = std::make_unique<QLinearB200Impl>(inter_, h, false, qa, options);
  }
  final_norm_ = std::make_unique<RMSNormImpl>(h, args.rms_norm_eps, options);
}

void LlamaDecoderStep::prepack() {
  for (auto& l : layers_)
    for (QLinearB200Impl* m : {l.qkv.get(), l.o.get(), l.gate_up.get(), l.down.get()})
      if (m != nullptr) m->ensure_packed();
}

void LlamaDecoderStep::load_state_dict(const StateDict& sd) {
  auto sub = [&](const std::string& prefix) {
    StateDict out;
    for (const auto& kv : sd)
      if (kv.first.rfind(prefix, 0) == 0) out.emplace(kv.first.substr(prefix.size()), kv.second);
    return out;
  };
  const int64_t w = parallel_args_.world_size(), r = parallel_args_.rank();
  for (size_t i = 0; i < layers_.size(); ++i) {
    const std::string p = "layers." + std::to_string(i) + ".";
    if (w == 1) {
      layers_[i].qkv->load_state_dict(sub(p + "qkv."));
      layers_[i].o->load_state_dict(sub(p + "o."));
      layers_[i].gate_up->load_state_dict(sub(p + "gate_up."));
      layers_[i].down->load_state_dict(sub(p + "down."));
    } else {
      const LlamaLayerShards sh = shard_llama_layer(sub(p + "qkv."), sub(p + "o."), sub(p + "gate_up."),
                                                    sub(p + "down."), args_, quant_args_,
                                                    static_cast<int>(r), static_cast<int>(w));
      layers_[i].qkv->load_state_dict(sh.qkv);
      layers_[i].o->load_state_dict(sh.o);
      layers_[i].gate_up->load_state_dict(sh.gate_up);
      layers_[i].down->load_state_dict(sh.down);
    }
    layers_[i].input_norm->load_state_dict(sub(p + "input_norm."));
    layers_[i].post_norm->load_state_dict(sub(p + "post_norm."));
  }
  final_norm_->load_state_dict(sub("final_norm."));
  // ParallelEmbedding: split on the hidden dim + all-gather (embedding.h:74-79); lm_head: column
  // parallel on the vocabulary + all-gather of the logits (llama.h:281-289)
  const int64_t hs = args_.hidden_size / w, vs = args_.vocab_size / w;
  embed_ = need(sd, "embed.weight").slice(1, r * hs, (r + 1) * hs).to(options_).contiguous();
  lm_head_ = need(sd, "lm_head.weight").slice(0, r * vs, (r + 1) * vs).to(options_).contiguous();
}

torch::Tensor LlamaDecoderStep::forward(const torch::Tensor& tokens, const torch::Tensor& positions,
                                        const InputParameters& params) {
  TORCH_CHECK(kv_caches_.size() == layers_.size(), "set_kv_caches() first");
  const int64_t H = n_heads_, Hkv = n_kv_heads_, D = args_.head_dim;  // this rank's heads
  const int64_t q_size = H * D, kv_size = Hkv * D;
  const int64_t w = parallel_args_.world_size();
  torch::Tensor h = embed_.index_select(0, tokens.to(torch::kLong));  // residual stream [T, hidden]
  if (w > 1) h = pg_->allgather_lastdim(h);
  const int64_t T = h.size(0);
  const auto dtype = h.scalar_type();

  // `pending` = output of the previous block, not yet added to the residual stream: a bf16 tensor
  // or the producing GEMM's partials (llama.h:170-177 with the adds folded into the norms).  Under
  // tensor parallelism the pending tensor is a row-parallel linear's local result: its all-reduce
  // happens here, fused with the GEMM's reduction — and with the add + norm when the shape allows.
  bool have_pending = false, pending_is_partials = false;
  torch::Tensor pending;
  W4Partials pending_parts;
  auto norm_residual = [&](RMSNormImpl& norm) {
    if (!pending_is_partials) {
      if (w > 1) pg_->allreduce(pending);
      return norm.forward_residual(pending, h);
    }
    if (w == 1) return norm.forward_residual_partials(pending_parts, h);
    if (pg_->supports_partials_norm(T, h.size(1), dtype))
      return pg_->allreduce_partials_norm(pending_parts.data, pending_parts.K, h, norm.weight, norm.eps());
    torch::Tensor reduced = pg_->allreduce_partials(pending_parts.data, pending_parts.K, dtype);
    return norm.forward_residual(reduced, h);
  };

  for (size_t li = 0; li < layers_.size(); ++li) {
    Layer& L = layers_[li];
    KVCache& cache = kv_caches_[li];
    torch::Tensor n1 = have_pending ? norm_residual(*L.input_norm) : L.input_norm->forward(h);

    torch::Tensor q, attn = torch::empty({T, H, D}, h.options());
    if (fuse_partials && L.qkv->supports_partials(T) && handler_->has_rope() && !cache.empty()) {
      torch::Tensor qkv = handler_->qkv_from_partials(L.qkv->forward_partials(n1), H, Hkv, D,
                                                      positions, cache, params, dtype);
      q = qkv.slice(1, 0, q_size).view({T, H, D});
    } else {
      torch::Tensor qkv = L.qkv->forward(n1);
      q = qkv.slice(1, 0, q_size).view({T, H, D});
      torch::Tensor k = qkv.slice(1, q_size, q_size + kv_size).view({T, Hkv, D});
      torch::Tensor v = qkv.slice(1, q_size + kv_size, q_size + 2 * kv_size).view({T, Hkv, D});
      handler_->apply_pos_emb_and_append(q, k, v, positions, cache, params);
    }
    handler_->batch_decode(q, cache, params, /*sliding_window=*/-1, attn);
    torch::Tensor attn2 = attn.view({T, q_size});

    // h += o_proj(attn); n2 = post_norm(h)
    pending_is_partials = fuse_partials && L.o->supports_partials(T);
    if (pending_is_partials) pending_parts = L.o->forward_partials(attn2);
    else pending = L.o->forward(attn2);
    torch::Tensor n2 = norm_residual(*L.post_norm);

    torch::Tensor act;
    if (fuse_partials && L.gate_up->supports_partials(T)) {
      act = kernel::silu_mul_partials(L.gate_up->forward_partials(n2), dtype);
    } else {
      act = kernel::silu_with_mul(L.gate_up->forward(n2));
    }
    pending_is_partials = fuse_partials && L.down->supports_partials(T);
    if (pending_is_partials) pending_parts = L.down->forward_partials(act);
    else pending = L.down->forward(act);
    have_pending = true;
  }
  torch::Tensor hn = have_pending ? norm_residual(*final_norm_) : final_norm_->forward(h);
  // dense bf16 lm_head: csrc/dense.cu where its shape allows (K % 64 == 0, N % 8 == 0), else the library GEMM
  torch::Tensor logits = (lm_head_.size(1) % 64 == 0 && lm_head_.size(0) % 8 == 0)
                             ? kernel::dense_linear(hn, lm_head_)
                             : torch::linear(hn, lm_head_);
  return w > 1 ? pg_->allgather_lastdim(logits) : logits;
}

torch::Tensor LlamaDecoderStep::step(const torch::Tensor& tokens, const torch::Tensor& positions,
                                     const InputParameters& params) {
  return kernel::argmax(forward(tokens, positions, params));
}

// ---------------------------------------------------------------------------------------------
// CUDA-graph replay of the step
// ---------------------------------------------------------------------------------------------
void CudaGraphStep::capture(LlamaDecoderStep* model, const torch::Tensor& tokens,
                            const torch::Tensor& positions, const InputParameters& params,
                            int64_t max_block_table_len, bool greedy) {
  TORCH_CHECK(graph_ == nullptr, "graph already captured");
  TORCH_CHECK(model != nullptr && tokens.is_cuda(), "CudaGraphStep: CUDA tensors only");
  TORCH_CHECK(max_block_table_len >= params.block_tables.size(0), "block table capacity too small");
  batch_size_ = params.num_sequences;
  num_tokens_ = tokens.size(0);
  // own the inputs: the graph reads these buffers, replay() refreshes them
  tokens_ = tokens.clone();
  positions_ = positions.clone();
  params_ = params;
  params_.q_cu_seq_lens = params.q_cu_seq_lens.clone();
  params_.kv_cu_seq_lens = params.kv_cu_seq_lens.clone();
  params_.new_cache_slots = params.new_cache_slots.clone();
  params_.cu_block_lens = params.cu_block_lens.clone();
  params_.block_tables = torch::zeros({max_block_table_len}, params.block_tables.options());
  params_.block_tables.slice(0, 0, params.block_tables.size(0)).copy_(params.block_tables);

  auto run = [&]() {
    return greedy ? model->step(tokens_, positions_, params_) : model->forward(tokens_, positions_, params_);
  };
  // warm up outside the capture: workspaces, tensor maps, library handles
  torch::cuda::synchronize();
  run();
  torch::cuda::synchronize();
  {
    at::cuda::CUDAStream capture_stream = at::cuda::getStreamFromPool();
    at::cuda::CUDAStreamGuard stream_guard(capture_stream);
    graph_ = std::make_unique<at::cuda::CUDAGraph>();
    graph_->capture_begin(at::cuda::graph_pool_handle(), cudaStreamCaptureModeThreadLocal);
    output_ = run();
    graph_->capture_end();
  }
  torch::cuda::synchronize();
}

torch::Tensor CudaGraphStep::replay(const torch::Tensor& tokens, const torch::Tensor& positions,
                                    const InputParameters& params) {
  TORCH_CHECK(graph_ != nullptr, "graph not captured");
  TORCH_CHECK(params.num_sequences == batch_size_, "batch size mismatch");
  TORCH_CHECK(tokens.size(0) == num_tokens_, "num tokens mismatch");
  const int64_t table_len = params.block_tables.size(0);
  TORCH_CHECK(params_.block_tables.size(0) >= table_len, "block table size mismatch");
  TORCH_CHECK(params.kv_max_seq_len <= params_.kv_max_seq_len && params.q_max_seq_len <= params_.q_max_seq_len,
              "step exceeds the sequence lengths the graph was captured for");
  tokens_.copy_(tokens, /*non_blocking=*/true);
  positions_.copy_(positions, /*non_blocking=*/true);
  params_.q_cu_seq_lens.copy_(params.q_cu_seq_lens, /*non_blocking=*/true);
  params_.kv_cu_seq_lens.copy_(params.kv_cu_seq_lens, /*non_blocking=*/true);
  params_.new_cache_slots.copy_(params.new_cache_slots, /*non_blocking=*/true);
  // the block table may arrive with a different padding length
  params_.block_tables.slice(0, 0, table_len).copy_(params.block_tables, /*non_blocking=*/true);
  params_.cu_block_lens.copy_(params.cu_block_lens, /*non_blocking=*/true);
  graph_->replay();
  return output_;
}

// ---------------------------------------------------------------------------------------------
// graph per batch size, eager otherwise
// ---------------------------------------------------------------------------------------------
void ModelRunner::capture_cuda_graphs(uint32_t batch_size) {
  TORCH_CHECK(device_.is_cuda(), "CUDA graphs need a CUDA device");
  TORCH_CHECK(graphs_.find(batch_size) == graphs_.end(), "batch size ", batch_size, " already captured");
  c10::cuda::CUDAGuard guard(device_);
  const int64_t ndt = options_.num_decoding_tokens, bs = options_.block_size;
  const int64_t n_tokens = ndt * batch_size;
  const auto i32 = torch::dtype(torch::kInt32).device(device_);
  // round up and add one block per sequence (speculative decoding), model_runner.cpp:57-60
  const int64_t max_block_table_len = (options_.cuda_graph_max_seq_len + bs - 1) / bs + 1;
  InputParameters params;
  params.num_sequences = static_cast<int32_t>(batch_size);
  params.q_max_seq_len = static_cast<int32_t>(ndt);
  params.kv_max_seq_len = static_cast<int32_t>(options_.cuda_graph_max_seq_len);
  params.q_cu_seq_lens = torch::arange(0, n_tokens + 1, ndt, i32);
  params.kv_cu_seq_lens = torch::arange(0, n_tokens + 1, ndt, i32);
  // placeholder sequences of kv_len = ndt, all in block 0 (first-slot id 0)
  const int64_t nb = (ndt + bs - 1) / bs;
  params.new_cache_slots = torch::arange(0, n_tokens, i32).remainder(std::min(ndt, bs));
  params.block_tables = torch::zeros({static_cast<int64_t>(batch_size) * nb}, i32);
  params.cu_block_lens = torch::arange(0, static_cast<int64_t>(batch_size) * nb + 1, nb, i32);
  auto graph = std::make_unique<CudaGraphStep>();
  graph->capture(model_, torch::zeros({n_tokens}, i32), torch::zeros({n_tokens}, i32), params,
                 static_cast<int64_t>(batch_size) * max_block_table_len, options_.greedy);
  graphs_[batch_size] = std::move(graph);
}

torch::Tensor ModelRunner::forward(const torch::Tensor& tokens, const torch::Tensor& positions,
                                   const InputParameters& params) {
  const uint32_t batch_size = static_cast<uint32_t>(params.num_sequences);
  auto it = graphs_.find(batch_size);
  if (it != graphs_.end()) {
    const bool seq_len_supported = params.kv_max_seq_len <= options_.cuda_graph_max_seq_len;
    const bool same_num_decoding_tokens =
        params.q_max_seq_len == options_.num_decoding_tokens &&
        tokens.size(0) == static_cast<int64_t>(batch_size) * options_.num_decoding_tokens;
    if (seq_len_supported && same_num_decoding_tokens) {
      ++n_replayed_;
      // the graph was captured for the longest context: replay with those host scalars
      InputParameters p = params;
      p.kv_max_seq_len = static_cast<int32_t>(options_.cuda_graph_max_seq_len);
      return it->second->replay(tokens, positions, p);
    }
  }
  ++n_eager_;
  return options_.greedy ? model_->step(tokens, positions, params) : model_->forward(tokens, positions, params);
}

}  // namespace llm
