// Python face of the reference's OWN decode kernels, compiled from /root/reference where they lie
// (oracle/ref/Makefile; TEST INFRASTRUCTURE, never on the product path).  Nothing of the reference
// is copied here: this file only includes its headers and re-exports the functions they declare.
//
//   src/kernels/layernorm_kernels.h:6-26     rms_norm, rms_norm_residual
//   src/kernels/pos_embedding_kernels.h:7-13 apply_rotary_pos_emb
//   src/kernels/kv_cache_kernels.h:6-11      set_kv_cache
//   src/kernels/activation_kernels.h:6-14    silu, silu_with_mul
//   src/kernels/attention/attn_api.h:12-27   paged_kv_varlen_mha (bf16 / fp16, head_dim 128 only:
//                                            the other head dims are not instantiated, so the module
//                                            must be imported with RTLD_LAZY)
//   src/kernels/quantization/marlin.h:17-28  marlin_gemm = marlin::gptq_gemm (4-bit, group 128,
//                                            17 <= M <= 64, N % 256 == 0: the instantiations built)
//   marlin_dequant_table                     known answers of the Marlin int4 -> bf16 weight
//                                            arithmetic (marlin_dequant_kat.cu around
//                                            quantization/marlin/numeric_conversion.h)
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include "attn_api.h"
#include "marlin.h"

extern "C" int marlin_dequant_kat(const void* scales, int S, void* out_zp, void* out_sym, void* stream);
#include "activation_kernels.h"
#include "kv_cache_kernels.h"
#include "layernorm_kernels.h"
#include "pos_embedding_kernels.h"

// src/kernels/sampling/sampling_kernels.h:7-29 (declared here: that header also pulls in curand for
// the top-k sampler, which is not part of this build)
namespace llm::kernel {
void apply_temperature_penalty(torch::Tensor& logits, const torch::Tensor& temperatures);
void apply_repetition_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                              const torch::Tensor& token_ids_lens, const torch::Tensor& penalities);
void apply_frequency_presence_penalty(torch::Tensor& logits, const torch::Tensor& token_ids,
                                      const torch::Tensor& token_counts, const torch::Tensor& token_ids_lens,
                                      const torch::Tensor& frequency_penalties,
                                      const torch::Tensor& presence_penalties);
void invoke_softmax(torch::Tensor& logits);
}  // namespace llm::kernel

PYBIND11_MODULE(_ref_kernels, m) {
  m.doc() = "vectorch-ai/ScaleLLM src/kernels compiled for sm_100a (test oracle)";
  m.def("rms_norm", [](torch::Tensor out, torch::Tensor x, torch::Tensor w, double eps) {
    llm::kernel::rms_norm(out, x, w, static_cast<float>(eps));
  });
  m.def("rms_norm_residual",
        [](torch::Tensor out, torch::Tensor res, torch::Tensor x, torch::Tensor w, double eps) {
          llm::kernel::rms_norm_residual(out, res, x, w, static_cast<float>(eps));
        });
  m.def("apply_rotary_pos_emb", [](torch::Tensor q, torch::Tensor k, torch::Tensor pos,
                                   torch::Tensor cos_sin, int rotary_dim, bool interleaved) {
    llm::kernel::apply_rotary_pos_emb(q, k, pos, cos_sin, rotary_dim, interleaved);
  });
  m.def("set_kv_cache", [](torch::Tensor slots, torch::Tensor k, torch::Tensor v, torch::Tensor kc,
                           torch::Tensor vc) { llm::kernel::set_kv_cache(slots, k, v, kc, vc); });
  m.def("paged_kv_varlen_mha",
        [](torch::Tensor out, torch::Tensor q, torch::Tensor kc, torch::Tensor vc, torch::Tensor q_cu,
           torch::Tensor kv_cu, torch::Tensor table, torch::Tensor blk_cu,
           std::optional<torch::Tensor> alibi, int bs, int max_q, int max_kv, double scale, double cap,
           int window) {
          TORCH_CHECK(q.size(-1) == 128, "reference attention was instantiated for head_dim 128 only");
          llm::paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, alibi, bs, max_q, max_kv,
                                   static_cast<float>(scale), static_cast<float>(cap), window);
        });
  m.def("marlin_gemm", [](torch::Tensor A, torch::Tensor B, torch::Tensor C, torch::Tensor scales,
                          torch::Tensor zeros, torch::Tensor g_idx, torch::Tensor perm,
                          torch::Tensor workspace, int num_bits, bool is_k_full, bool has_zp,
                          bool use_fp32_reduce) {
    marlin::gptq_gemm(A, B, C, scales, zeros, g_idx, perm, workspace, num_bits, is_k_full, has_zp,
                      use_fp32_reduce);
  });
  m.def("marlin_dequant_table", [](torch::Tensor scales) {
    TORCH_CHECK(scales.is_cuda() && scales.scalar_type() == torch::kBFloat16 && scales.is_contiguous());
    const int S = static_cast<int>(scales.numel());
    torch::Tensor zp = torch::empty({16, 16, S}, scales.options());   // [q][z][s]
    torch::Tensor sym = torch::empty({16, S}, scales.options());      // [q][s], zero point 8 built in
    const int rc = marlin_dequant_kat(scales.const_data_ptr(), S, zp.data_ptr(), sym.data_ptr(),
                                      at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == 0, "marlin_dequant_kat launch failed: ", rc);
    return std::make_tuple(zp, sym);
  });
  m.def("gemma_rms_norm", [](torch::Tensor out, torch::Tensor x, torch::Tensor w, double eps) {
    llm::kernel::gemma_rms_norm(out, x, w, static_cast<float>(eps));
  });
  m.def("layer_norm", [](torch::Tensor out, torch::Tensor x, torch::Tensor w,
                         std::optional<torch::Tensor> b, double eps) {
    llm::kernel::layer_norm(out, x, w, b.has_value() ? *b : torch::Tensor(), static_cast<float>(eps));
  });
  m.def("gelu_new", &llm::kernel::gelu_new);
  m.def("gelu_fast", &llm::kernel::gelu_fast);
  m.def("gelu_new_with_mul", &llm::kernel::gelu_new_with_mul);
  m.def("gelu_fast_with_mul", &llm::kernel::gelu_fast_with_mul);
  m.def("apply_temperature_penalty", [](torch::Tensor logits, torch::Tensor t) {
    llm::kernel::apply_temperature_penalty(logits, t);
  });
  m.def("apply_repetition_penalty", [](torch::Tensor logits, torch::Tensor ids, torch::Tensor lens, torch::Tensor p) {
    llm::kernel::apply_repetition_penalty(logits, ids, lens, p);
  });
  m.def("apply_frequency_presence_penalty", [](torch::Tensor logits, torch::Tensor ids, torch::Tensor counts,
                                               torch::Tensor lens, torch::Tensor f, torch::Tensor p) {
    llm::kernel::apply_frequency_presence_penalty(logits, ids, counts, lens, f, p);
  });
  m.def("invoke_softmax", [](torch::Tensor logits) { llm::kernel::invoke_softmax(logits); });
  m.def("silu", &llm::kernel::silu);
  m.def("silu_with_mul", &llm::kernel::silu_with_mul);
}
