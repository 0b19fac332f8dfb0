// Python face of the C++ shim (the counterpart of scalellm/csrc/kernels.cu:9-55) so the GPU tests
// can call the reference-signature C++ functions directly.
#include <torch/extension.h>

#include "b200_kernels.h"
#include "b200_layers.h"
#include "b200_process_group.h"

PYBIND11_MODULE(_b200_shim, m) {
  m.def("rms_norm", [](torch::Tensor out, torch::Tensor x, torch::Tensor w, double eps) {
    llm::kernel::rms_norm(out, x, w, static_cast<float>(eps));
  });
  m.def("rms_norm_residual",
        [](torch::Tensor out, torch::Tensor res, torch::Tensor x, torch::Tensor w, double eps) {
          llm::kernel::rms_norm_residual(out, res, x, w, static_cast<float>(eps));
        });
  m.def("apply_rotary_pos_emb", [](torch::Tensor q, torch::Tensor k, torch::Tensor pos,
                                   torch::Tensor cs, int rot, bool il) {
    llm::kernel::apply_rotary_pos_emb(q, k, pos, cs, rot, il);
  });
  m.def("set_kv_cache", [](torch::Tensor slots, torch::Tensor k, torch::Tensor v, torch::Tensor kc,
                           torch::Tensor vc) { llm::kernel::set_kv_cache(slots, k, v, kc, vc); });
  m.def("apply_top_k_top_p", [](torch::Tensor logits, c10::optional<torch::Tensor> top_k,
                                c10::optional<torch::Tensor> top_p) {
    llm::kernel::apply_top_k_top_p(logits, top_k.value_or(torch::Tensor()), top_p.value_or(torch::Tensor()));
  });
  m.def("silu", &llm::kernel::silu);
  m.def("silu_with_mul", &llm::kernel::silu_with_mul);
  m.def("paged_kv_varlen_mha",
        [](torch::Tensor out, torch::Tensor q, torch::Tensor kc, torch::Tensor vc,
           torch::Tensor q_cu, torch::Tensor kv_cu, torch::Tensor table, torch::Tensor blk_cu,
           std::optional<torch::Tensor> alibi, int bs, int max_q, int max_kv, double scale,
           double cap, int window) {
          llm::paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, alibi, bs, max_q,
                                   max_kv, static_cast<float>(scale), static_cast<float>(cap),
                                   window);
        });
  m.def("marlin_awq_repack", [](torch::Tensor qw, torch::Tensor qz, torch::Tensor s,
                                torch::Tensor out, int64_t g) { marlin::b200_awq_repack(qw, qz, s, out, g); });
  m.def("marlin_gptq_repack", [](torch::Tensor qw, torch::Tensor s, torch::Tensor out, int64_t g) {
    marlin::b200_gptq_repack(qw, s, out, g);
  });
  m.def("marlin_gemm", [](torch::Tensor A, torch::Tensor B, torch::Tensor C, torch::Tensor scales,
                          torch::Tensor zeros, torch::Tensor g_idx, torch::Tensor perm,
                          torch::Tensor ws, int bits, bool k_full, bool has_zp, bool fp32r) {
    marlin::gptq_gemm(A, B, C, scales, zeros, g_idx, perm, ws, bits, k_full, has_zp, fp32r);
  });
  // the reference's exact signatures (scalellm/csrc/kernels.cu:9-55 exposes the same three)
  m.def("marlin_awq_repack_ref", [](torch::Tensor qw, torch::Tensor out, int64_t bits) {
    marlin::awq_repack(qw, out, bits);
  });
  m.def("marlin_gptq_repack_ref", [](torch::Tensor qw, torch::Tensor perm, torch::Tensor out, int64_t bits) {
    marlin::gptq_repack(qw, perm, out, bits);
  });
  m.def("assembled_weights", &marlin::b200_assembled_weights);
  m.def("packed_bytes", &marlin::b200_packed_bytes);
  m.def("workspace_bytes", &marlin::b200_workspace_bytes);

  namespace py = pybind11;
  // ---- tensor-parallel plumbing in the reference's threading model (shim/b200_process_group.h) --
  py::class_<llm::ProcessGroup>(m, "ProcessGroup")
      .def("rank", &llm::ProcessGroup::rank)
      .def("world_size", &llm::ProcessGroup::world_size)
      .def("allreduce", [](const llm::ProcessGroup& self, torch::Tensor t) { self.allreduce(t); })
      .def("allgather",
           [](const llm::ProcessGroup& self, torch::Tensor in, std::vector<torch::Tensor> outs) {
             self.allgather(in, outs);
           })
      .def("gather_from_model_parallel_region",
           [](llm::ProcessGroup& self, torch::Tensor in) {
             return llm::gather_from_model_parallel_region(
                 in, llm::ParallelArgs(self.rank(), self.world_size(), &self));
           })
      .def("reduce_from_model_parallel_region",
           [](llm::ProcessGroup& self, torch::Tensor in) {
             return llm::reduce_from_model_parallel_region(
                 in, llm::ParallelArgs(self.rank(), self.world_size(), &self));
           })
      .def("scatter_to_model_parallel_region", [](llm::ProcessGroup& self, torch::Tensor in) {
        return llm::scatter_to_model_parallel_region(
            in, llm::ParallelArgs(self.rank(), self.world_size(), &self));
      });
  m.def("create_process_groups", [](const std::vector<int>& device_indices) {
    std::vector<torch::Device> devices;
    for (int i : device_indices) devices.emplace_back(torch::kCUDA, static_cast<c10::DeviceIndex>(i));
    return llm::ProcessGroup::create_process_groups(devices);
  });

  // ---- plugin-level host side (shim/b200_layers.h) -------------------------------------------
  py::class_<llm::LlamaDecoderStep>(m, "LlamaDecoderStep")
      .def(py::init([](int64_t hidden, int64_t n_layers, int64_t n_heads, int64_t n_kv_heads,
                       int64_t head_dim, int64_t inter, int64_t vocab, int64_t max_pos, double eps,
                       std::string quant_method, int64_t group_size, bool is_sym,
                       torch::Tensor inv_freq, torch::Tensor like, llm::ProcessGroup* pg) {
        llm::LlamaArgs a;
        a.hidden_size = hidden;
        a.n_layers = n_layers;
        a.n_heads = n_heads;
        a.n_kv_heads = n_kv_heads;
        a.head_dim = head_dim;
        a.intermediate_size = inter;
        a.vocab_size = vocab;
        a.max_position_embeddings = max_pos;
        a.rms_norm_eps = static_cast<float>(eps);
        llm::QuantArgs q;
        q.quant_method = std::move(quant_method);
        q.group_size = group_size;
        q.is_sym = is_sym;
        const llm::ParallelArgs pa(pg ? pg->rank() : 0, pg ? pg->world_size() : 1, pg);
        return std::make_unique<llm::LlamaDecoderStep>(a, q, inv_freq, like.options(), pa);
      }),
           py::arg("hidden"), py::arg("n_layers"), py::arg("n_heads"), py::arg("n_kv_heads"),
           py::arg("head_dim"), py::arg("inter"), py::arg("vocab"), py::arg("max_pos"), py::arg("eps"),
           py::arg("quant_method"), py::arg("group_size"), py::arg("is_sym"), py::arg("inv_freq"),
           py::arg("like"), py::arg("process_group") = nullptr, py::keep_alive<1, 16>())
      .def("load_state_dict",
           [](llm::LlamaDecoderStep& self, const std::unordered_map<std::string, torch::Tensor>& sd) {
             self.load_state_dict(sd);
           })
      .def("set_kv_caches",
           [](llm::LlamaDecoderStep& self, const std::vector<torch::Tensor>& k,
              const std::vector<torch::Tensor>& v, int64_t block_size) {
             std::vector<llm::KVCache> caches;
             for (size_t i = 0; i < k.size(); ++i) caches.emplace_back(k[i], v[i], block_size);
             self.set_kv_caches(std::move(caches));
           })
      .def_readwrite("fuse_partials", &llm::LlamaDecoderStep::fuse_partials)
      .def("forward",
           [](llm::LlamaDecoderStep& self, torch::Tensor tokens, torch::Tensor positions,
              torch::Tensor q_cu, torch::Tensor kv_cu, int kv_max, int q_max, torch::Tensor slots,
              torch::Tensor tables, torch::Tensor blk_cu) {
             llm::InputParameters p;
             p.num_sequences = static_cast<int32_t>(q_cu.size(0) - 1);
             p.q_cu_seq_lens = q_cu;
             p.kv_cu_seq_lens = kv_cu;
             p.kv_max_seq_len = kv_max;
             p.q_max_seq_len = q_max;
             p.new_cache_slots = slots;
             p.block_tables = tables;
             p.cu_block_lens = blk_cu;
             py::gil_scoped_release release;  // ranks driven from Python threads run concurrently
             return self.forward(tokens, positions, p);
           });

  m.def("shard_llama_layer",
        [](const llm::StateDict& qkv, const llm::StateDict& o, const llm::StateDict& gate_up,
           const llm::StateDict& down, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim,
           int64_t inter, std::string quant_method, int64_t group_size, int rank, int world) {
          llm::LlamaArgs a;
          a.n_heads = n_heads;
          a.n_kv_heads = n_kv_heads;
          a.head_dim = head_dim;
          a.intermediate_size = inter;
          llm::QuantArgs q;
          q.quant_method = std::move(quant_method);
          q.group_size = group_size;
          const auto sh = llm::shard_llama_layer(qkv, o, gate_up, down, a, q, rank, world);
          return std::make_tuple(sh.qkv, sh.o, sh.gate_up, sh.down);
        });
  py::class_<llm::CudaGraphStep>(m, "CudaGraphStep")
      .def(py::init<>())
      .def("capture",
           [](llm::CudaGraphStep& self, llm::LlamaDecoderStep& model, torch::Tensor tokens,
              torch::Tensor positions, torch::Tensor q_cu, torch::Tensor kv_cu, int kv_max, int q_max,
              torch::Tensor slots, torch::Tensor tables, torch::Tensor blk_cu, int64_t max_table_len,
              bool greedy) {
             llm::InputParameters p;
             p.num_sequences = static_cast<int32_t>(q_cu.size(0) - 1);
             p.q_cu_seq_lens = q_cu;
             p.kv_cu_seq_lens = kv_cu;
             p.kv_max_seq_len = kv_max;
             p.q_max_seq_len = q_max;
             p.new_cache_slots = slots;
             p.block_tables = tables;
             p.cu_block_lens = blk_cu;
             self.capture(&model, tokens, positions, p, max_table_len, greedy);
           })
      .def("replay", [](llm::CudaGraphStep& self, torch::Tensor tokens, torch::Tensor positions,
                        torch::Tensor q_cu, torch::Tensor kv_cu, int kv_max, int q_max,
                        torch::Tensor slots, torch::Tensor tables, torch::Tensor blk_cu) {
        llm::InputParameters p;
        p.num_sequences = static_cast<int32_t>(q_cu.size(0) - 1);
        p.q_cu_seq_lens = q_cu;
        p.kv_cu_seq_lens = kv_cu;
        p.kv_max_seq_len = kv_max;
        p.q_max_seq_len = q_max;
        p.new_cache_slots = slots;
        p.block_tables = tables;
        p.cu_block_lens = blk_cu;
        return self.replay(tokens, positions, p);
      });



  py::class_<llm::ModelRunner>(m, "ModelRunner")
      .def(py::init([](llm::LlamaDecoderStep& model, int device_index, std::vector<uint32_t> batch_sizes,
                       int64_t num_decoding_tokens, int64_t max_seq_len, int64_t block_size, bool greedy) {
             llm::ModelRunner::Options o;
             o.cuda_graph_batch_sizes = std::move(batch_sizes);
             o.num_decoding_tokens = num_decoding_tokens;
             o.cuda_graph_max_seq_len = max_seq_len;
             o.block_size = block_size;
             o.greedy = greedy;
             return std::make_unique<llm::ModelRunner>(
                 &model, torch::Device(torch::kCUDA, static_cast<c10::DeviceIndex>(device_index)), o);
           }),
           py::keep_alive<1, 2>())
      .def("capture_cuda_graphs", &llm::ModelRunner::capture_cuda_graphs)
      .def("num_cuda_graph_replayed", &llm::ModelRunner::num_cuda_graph_replayed)
      .def("num_eager_execution", &llm::ModelRunner::num_eager_execution)
      .def("forward", [](llm::ModelRunner& self, torch::Tensor tokens, torch::Tensor positions,
                         torch::Tensor q_cu, torch::Tensor kv_cu, int kv_max, int q_max,
                         torch::Tensor slots, torch::Tensor tables, torch::Tensor blk_cu) {
        llm::InputParameters p;
        p.num_sequences = static_cast<int32_t>(q_cu.size(0) - 1);
        p.q_cu_seq_lens = q_cu;
        p.kv_cu_seq_lens = kv_cu;
        p.kv_max_seq_len = kv_max;
        p.q_max_seq_len = q_max;
        p.new_cache_slots = slots;
        p.block_tables = tables;
        p.cu_block_lens = blk_cu;
        return self.forward(tokens, positions, p);
      });
}
