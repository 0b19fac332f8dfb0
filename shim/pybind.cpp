// Python face of the C++ shim (the counterpart of scalellm/csrc/kernels.cu:9-55) so the GPU tests
// can call the reference-signature C++ functions directly.
#include <torch/extension.h>

#include "b200_kernels.h"

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
                                torch::Tensor out, int64_t g) { marlin::awq_repack(qw, qz, s, out, g); });
  m.def("marlin_gptq_repack", [](torch::Tensor qw, torch::Tensor s, torch::Tensor out, int64_t g) {
    marlin::gptq_repack(qw, s, out, g);
  });
  m.def("marlin_gemm", [](torch::Tensor A, torch::Tensor B, torch::Tensor C, torch::Tensor scales,
                          torch::Tensor zeros, torch::Tensor g_idx, torch::Tensor perm,
                          torch::Tensor ws, int bits, bool k_full, bool has_zp, bool fp32r) {
    marlin::gptq_gemm(A, B, C, scales, zeros, g_idx, perm, ws, bits, k_full, has_zp, fp32r);
  });
  m.def("packed_bytes", &marlin::b200_packed_bytes);
  m.def("workspace_bytes", &marlin::b200_workspace_bytes);
}
