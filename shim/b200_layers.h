// b200_layers.h — C++ host side of the decode step above the C ABI: the reference's plugin-level
// interfaces (same names, members and call order) implemented over libb200decode, plus the Llama
// decoder wiring that uses the fused paths.
//
//   InputParameters                      src/models/parameters.h:11-56
//   KVCache                              src/memory/kv_cache.h / kv_cache.cpp:15-98
//   AttentionHandler (+ B200Handler)     src/layers/attention/handler.h:15-65
//   ParallelLinearImpl (+ QLinearB200Impl)  src/layers/linear/parallel_linear.h:17-40,
//                                        src/layers/quantization/qlinear_awq_marlin_impl.cpp:129-365
//   RMSNormImpl                          src/layers/normalization.h:114-139
//   LlamaDecoderStep                     src/models/meta/llama.h:61-64,123-133,170-177,220-232,281-289
//   CudaGraphStep                        src/engine/model_runner.cpp:141-210 (ModelRunner::CudaGraph)
//   ModelRunner                          src/engine/model_runner.cpp:25-139 (graph per batch size)
//
// Inside ScaleLLM these classes derive from the engine's own headers; here the
// interfaces are restated so that the library is self-contained and testable.  Tensor parallelism
// follows the reference's partitioning (column split q/k/v/gate/up, row split o/down, one exchange
// per row-parallel linear, llama.h:83-115) over shim/b200_process_group.h: one LlamaDecoderStep per
// rank, all ranks in one process with one thread each, or one process per rank.
#pragma once

#include <torch/torch.h>

#include <ATen/cuda/CUDAGraph.h>

#include "b200_process_group.h"

#include <memory>
#include <optional>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace llm {

// ---------------------------------------------------------------------------------------------
// fused operator-level entry points that have no counterpart in the reference's kernel API
// ---------------------------------------------------------------------------------------------
// fp32 stream-K partials of a W4A16 GEMM (include/b200_decode.h "partials mode"); K = the GEMM's
// reduction dimension (a consumer recomputes the tile -> slot partition from (N, K, rows)).
struct W4Partials {
  torch::Tensor data;  // [slots, rows, N] fp32
  int64_t K = 0;
};

namespace kernel {
W4Partials w4a16_gemm_partials(const torch::Tensor& a, const torch::Tensor& packed, int64_t N,
                               int64_t group_size);
torch::Tensor w4a16_reduce_partials(const W4Partials& p, const std::optional<torch::Tensor>& bias);
// residual += T(sum partials); returns rms_norm(residual) * weight
torch::Tensor rms_norm_residual_partials(const W4Partials& p, torch::Tensor& residual,
                                         const torch::Tensor& weight, float eps);
// qkv partials -> qkv [T, (H + 2 Hkv) D] (q, k rotated), rotated k and v written to their slots
torch::Tensor rope_and_set_kv_cache_partials(const W4Partials& p, int64_t n_heads,
                                             int64_t n_kv_heads, int64_t head_dim,
                                             const torch::Tensor& positions,
                                             const torch::Tensor& cos_sin,
                                             const torch::Tensor& slot_ids, torch::Tensor& key_cache,
                                             torch::Tensor& value_cache, int rotary_dim,
                                             bool interleaved, torch::ScalarType dtype);
// gate_up partials [slots, rows, 2 I] -> silu(gate) * up [rows, I]
torch::Tensor silu_mul_partials(const W4Partials& p, torch::ScalarType dtype);
torch::Tensor dense_linear(const torch::Tensor& x, const torch::Tensor& w);  // csrc/dense.cu
torch::Tensor argmax(const torch::Tensor& logits);
}  // namespace kernel

// ---------------------------------------------------------------------------------------------
// step metadata and KV cache
// ---------------------------------------------------------------------------------------------
struct InputParameters {
  int32_t num_sequences = 0;
  torch::Tensor q_cu_seq_lens;    // int32 [n_seq + 1]
  torch::Tensor kv_cu_seq_lens;   // int32 [n_seq + 1]
  int32_t kv_max_seq_len = 0;
  int32_t q_max_seq_len = 0;
  torch::Tensor new_cache_slots;  // int32 [n_tokens]
  torch::Tensor block_tables;     // int32 [n_blocks]: first-slot ids (engine/batch.cpp:206-209)
  torch::Tensor cu_block_lens;    // int32 [n_seq + 1]
};

class KVCache {
 public:
  KVCache() = default;
  KVCache(torch::Tensor key_cache, torch::Tensor value_cache, int64_t block_size)
      : key_cache_(std::move(key_cache)), value_cache_(std::move(value_cache)),
        block_size_(block_size) {}
  bool empty() const { return !key_cache_.defined() || key_cache_.numel() == 0; }
  int64_t block_size() const { return block_size_; }
  std::tuple<torch::Tensor, torch::Tensor> get_kv_cache() const { return {key_cache_, value_cache_}; }
  void set_kv_cache(const torch::Tensor& slot_ids, const torch::Tensor& keys,
                    const torch::Tensor& values);

 private:
  torch::Tensor key_cache_, value_cache_;  // [n_slots, n_kv_heads, head_dim]
  int64_t block_size_ = 0;
};

// ---------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------
class AttentionHandler {
 public:
  virtual ~AttentionHandler() = default;
  virtual int64_t get_estimate_workspace_size() { return -1; }
  virtual void set_workspace(const torch::Tensor& /*workspace*/) {}
  virtual std::tuple<torch::Tensor, torch::Tensor> apply_pos_emb(const torch::Tensor& query,
                                                                 const torch::Tensor& key,
                                                                 const torch::Tensor& positions) = 0;
  virtual void batch_decode(const torch::Tensor& query, const KVCache& kv_cache,
                            const InputParameters& input_params, int32_t sliding_window,
                            torch::Tensor& output) = 0;
  virtual void append_kv_cache(KVCache& kv_cache, const torch::Tensor& key,
                               const torch::Tensor& value, const InputParameters& input_params) = 0;
};

class B200Handler final : public AttentionHandler {
 public:
  // cos_sin: [max_position, rotary_dim] in the model dtype (pos_embedding.cpp:183-215)
  B200Handler(float sm_scale, float logits_soft_cap, std::optional<torch::Tensor> alibi_slopes,
              torch::Tensor cos_sin, int64_t rotary_dim, bool interleaved);

  std::tuple<torch::Tensor, torch::Tensor> apply_pos_emb(const torch::Tensor& query,
                                                         const torch::Tensor& key,
                                                         const torch::Tensor& positions) override;
  void batch_decode(const torch::Tensor& query, const KVCache& kv_cache,
                    const InputParameters& input_params, int32_t sliding_window,
                    torch::Tensor& output) override;
  void append_kv_cache(KVCache& kv_cache, const torch::Tensor& key, const torch::Tensor& value,
                       const InputParameters& input_params) override;

  // B200 extensions: RoPE + KV-slot write in one launch; the same fed by the qkv GEMM's partials
  void apply_pos_emb_and_append(torch::Tensor& query, torch::Tensor& key, const torch::Tensor& value,
                                const torch::Tensor& positions, KVCache& kv_cache,
                                const InputParameters& input_params);
  torch::Tensor qkv_from_partials(const W4Partials& qkv, int64_t n_heads, int64_t n_kv_heads,
                                  int64_t head_dim, const torch::Tensor& positions,
                                  KVCache& kv_cache, const InputParameters& input_params,
                                  torch::ScalarType dtype);
  bool has_rope() const { return cos_sin_.defined(); }

 private:
  float sm_scale_, soft_cap_;
  std::optional<torch::Tensor> alibi_;
  torch::Tensor cos_sin_;
  int64_t rotary_dim_;
  bool interleaved_;
};

// ---------------------------------------------------------------------------------------------
// linear / norm layers
// ---------------------------------------------------------------------------------------------
struct QuantArgs {  // quant_args.h:10-33
  std::string quant_method;  // "awq" | "gptq"
  int64_t bits = 4;
  int64_t group_size = 128;
  bool desc_act = false;
  bool is_sym = false;
};

using StateDict = std::unordered_map<std::string, torch::Tensor>;

class ParallelLinearImpl {
 public:
  virtual ~ParallelLinearImpl() = default;
  virtual torch::Tensor forward(torch::Tensor input) = 0;
  virtual void load_state_dict(const StateDict& state_dict) = 0;
  virtual void verify_loaded_weights() const = 0;
};

// int4 linear of one rank: checkpoint tensors in, lazy repack on first use
// (qlinear_awq_marlin_impl.cpp:99-125,232-235), forward through the W4A16 GEMM.
class QLinearB200Impl final : public ParallelLinearImpl {
 public:
  QLinearB200Impl(int64_t in_features, int64_t out_features, bool bias, const QuantArgs& quant_args,
                  const torch::TensorOptions& options);
  torch::Tensor forward(torch::Tensor input) override;
  void load_state_dict(const StateDict& state_dict) override;  // qweight, qzeros (awq), scales, [bias]
  void verify_loaded_weights() const override;

  bool supports_partials(int64_t n_rows) const { return !bias_.defined() && n_rows > 0 && n_rows <= 128; }
  W4Partials forward_partials(const torch::Tensor& input);
  int64_t in_features() const { return K_; }
  int64_t out_features() const { return N_; }
  void ensure_packed();  // repack now (otherwise lazily at the first forward)

 private:
  int64_t K_, N_;
  QuantArgs qa_;
  torch::TensorOptions options_;
  bool has_bias_;
  torch::Tensor qweight_, qzeros_, scales_, bias_, packed_, workspace_;
  std::vector<torch::Tensor> retired_workspaces_;  // outgrown; captured graphs may still use them
};

// this rank's shard of a checkpoint-format int4 linear: output columns [c0, c1) of a column-
// parallel layer / input rows [k0, k1) of a row-parallel one (views, no copies)
StateDict shard_qlinear_columns(const StateDict& tensors, const QuantArgs& quant_args, int64_t c0,
                                int64_t c1);
StateDict shard_qlinear_rows(const StateDict& tensors, const QuantArgs& quant_args, int64_t k0,
                             int64_t k1);

class RMSNormImpl {
 public:
  RMSNormImpl(int64_t dim, float eps, const torch::TensorOptions& options);
  torch::Tensor forward(const torch::Tensor& input);
  torch::Tensor forward_residual(const torch::Tensor& input, torch::Tensor& residual);
  torch::Tensor forward_residual_partials(const W4Partials& input, torch::Tensor& residual);
  void load_state_dict(const StateDict& state_dict);
  float eps() const { return eps_; }
  torch::Tensor weight;

 private:
  float eps_;
};

// ---------------------------------------------------------------------------------------------
// the decode step
// ---------------------------------------------------------------------------------------------
struct LlamaArgs {
  int64_t hidden_size = 4096, n_layers = 32, n_heads = 32, n_kv_heads = 8, head_dim = 128;
  int64_t intermediate_size = 14336, vocab_size = 128256, max_position_embeddings = 8192;
  float rms_norm_eps = 1e-5f;
};

// One decoder layer's four int4 linears cut for `rank` of `world` from the unsharded fused tensors
// (qkv = q | k | v columns, gate_up = gate | up columns): llama.h:83-115 partitioning.
struct LlamaLayerShards {
  StateDict qkv, o, gate_up, down;
};
LlamaLayerShards shard_llama_layer(const StateDict& qkv, const StateDict& o, const StateDict& gate_up,
                                   const StateDict& down, const LlamaArgs& args,
                                   const QuantArgs& quant_args, int rank, int world);

class LlamaDecoderStep {
 public:
  // inv_freq: [head_dim / 2] fp32 (after any rope scaling, pos_embedding.cpp:75-109)
  LlamaDecoderStep(const LlamaArgs& args, const QuantArgs& quant_args, const torch::Tensor& inv_freq,
                   const torch::TensorOptions& options,
                   const ParallelArgs& parallel_args = ParallelArgs(0, 1, nullptr));

  // name -> tensor: "layers.<i>.{qkv,o,gate_up,down}.{qweight,qzeros,scales}",
  // "layers.<i>.{input_norm,post_norm}.weight", "embed.weight", "final_norm.weight", "lm_head.weight"
  // — the UNSHARDED tensors (qkv = q | k | v columns, gate_up = gate | up columns); each rank keeps
  // its shard: q/k/v/gate/up column ranges (kv heads replicated when n_kv_heads < world_size,
  // qkv_parallel_linear.cpp:28-70), o/down row ranges aligned to the quant groups
  // (qlinear_awq_marlin_impl.cpp:287), embedding split on hidden, lm_head on vocab.
  void load_state_dict(const StateDict& state_dict);
  void set_kv_caches(std::vector<KVCache> kv_caches) { kv_caches_ = std::move(kv_caches); }
  // Repack every int4 weight now (otherwise lazily at the first forward).  With all ranks in one
  // process (one thread per GPU) call this, and have the caching allocator hold enough memory,
  // BEFORE the first collective: a cudaMalloc by one rank synchronises with its peer-mapped
  // devices, so it dead-locks against another rank's all-reduce kernel spinning for it.
  void prepack();

  // logits [n_tokens, vocab]
  torch::Tensor forward(const torch::Tensor& tokens, const torch::Tensor& positions,
                        const InputParameters& params);
  // greedy next tokens [n_tokens] int64
  torch::Tensor step(const torch::Tensor& tokens, const torch::Tensor& positions,
                     const InputParameters& params);

  bool fuse_partials = true;  // GEMM reductions fused into their consumers (B200_FUSE_SPLITK)

 private:
  struct Layer {
    std::unique_ptr<RMSNormImpl> input_norm, post_norm;
    std::unique_ptr<QLinearB200Impl> qkv, o, gate_up, down;
  };
  LlamaArgs args_;
  QuantArgs quant_args_;
  torch::TensorOptions options_;
  ParallelArgs parallel_args_;
  ProcessGroupB200* pg_ = nullptr;          // the NVLink group, when world_size > 1
  int64_t n_heads_, n_kv_heads_, inter_;    // per rank
  std::vector<Layer> layers_;
  std::unique_ptr<RMSNormImpl> final_norm_;
  std::unique_ptr<B200Handler> handler_;
  torch::Tensor embed_, lm_head_;  // [vocab, h] each
  std::vector<KVCache> kv_caches_;
};

// One captured decode step for a fixed batch size and token count; replay() copies the step's
// metadata into the captured tensors and replays (ModelRunner::CudaGraph, model_runner.cpp:141-210).
// The host scalars of the captured launch (q_max_seq_len, kv_max_seq_len) are those of the capture:
// capture with kv_max_seq_len = the longest context replays will see (cuda_graph_max_seq_len).
class CudaGraphStep {
 public:
  // max_block_table_len: entries of the captured block table (model_runner.cpp:57-60:
  // batch * ((max_seq_len + block_size - 1) / block_size + 1)); greedy: return next tokens
  // (LlamaDecoderStep::step) instead of logits
  void capture(LlamaDecoderStep* model, const torch::Tensor& tokens, const torch::Tensor& positions,
               const InputParameters& params, int64_t max_block_table_len, bool greedy);
  // the returned tensor is the graph's output buffer: valid until the next replay
  torch::Tensor replay(const torch::Tensor& tokens, const torch::Tensor& positions,
                       const InputParameters& params);

 private:
  std::unique_ptr<at::cuda::CUDAGraph> graph_;
  int64_t batch_size_ = 0, num_tokens_ = 0;
  torch::Tensor tokens_, positions_;
  InputParameters params_;  // owns the captured metadata tensors
  torch::Tensor output_;
};

// Graph per captured batch size, eager otherwise (ModelRunner, model_runner.cpp:25-139): the
// selection rule of ModelRunner::forward — a captured graph is replayed when the batch size was
// captured, every sequence decodes num_decoding_tokens tokens and kv_max_seq_len fits
// cuda_graph_max_seq_len; anything else (prefill, odd batch sizes, long contexts) runs eagerly.
class ModelRunner {
 public:
  struct Options {
    std::vector<uint32_t> cuda_graph_batch_sizes;
    int64_t num_decoding_tokens = 1;
    int64_t cuda_graph_max_seq_len = 2048;
    int64_t block_size = 8;
    bool greedy = false;  // return next tokens instead of logits
  };
  ModelRunner(LlamaDecoderStep* model, const torch::Device& device, const Options& options)
      : model_(model), device_(device), options_(options) {}

  // Captures with placeholder metadata (every sequence at kv_len = num_decoding_tokens in block 0,
  // like model_runner.cpp:25-65): call it before the KV cache holds anything — the warm-up step
  // writes slot 0.
  void capture_cuda_graphs(uint32_t batch_size);
  torch::Tensor forward(const torch::Tensor& tokens, const torch::Tensor& positions,
                        const InputParameters& params);
  int64_t num_cuda_graph_replayed() const { return n_replayed_; }
  int64_t num_eager_execution() const { return n_eager_; }

 private:
  LlamaDecoderStep* model_;
  torch::Device device_;
  Options options_;
  std::unordered_map<uint32_t, std::unique_ptr<CudaGraphStep>> graphs_;
  int64_t n_replayed_ = 0, n_eager_ = 0;
};

}  // namespace llm
