// b200_process_group.cpp — see b200_process_group.h.  Host plumbing only; the collectives are
// csrc/allreduce.cu behind the C ABI.
#include "b200_process_group.h"

#include <algorithm>
#include <cstdlib>
#include <string>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "../include/b200_decode.h"

namespace llm {
namespace {

int dtype_code(const torch::Tensor& t) {
  switch (t.scalar_type()) {
    case torch::kBFloat16: return B200_BF16;
    case torch::kHalf: return B200_FP16;
    case torch::kFloat: return B200_FP32;
    default: TORCH_CHECK(false, "ProcessGroupB200: unsupported dtype ", t.scalar_type());
  }
}

void ok(int rc, const char* what) {
  TORCH_CHECK(rc == B200_OK, "b200 ", what, " failed (", rc, "): ", b200_last_error());
}

}  // namespace

std::vector<std::unique_ptr<ProcessGroup>> ProcessGroup::create_process_groups(
    const std::vector<torch::Device>& devices) {
  TORCH_CHECK(!devices.empty(), "devices should not be empty");
  const int world = static_cast<int>(devices.size());
  std::vector<int> idx;
  for (const auto& d : devices) {
    TORCH_CHECK(d.is_cuda(), "device should be cuda device");
    idx.push_back(d.index());
  }
  std::vector<b200_ar_comm*> comms(world, nullptr);
  ok(b200_ar_create_all(comms.data(), idx.data(), world, ProcessGroupB200::kMaxBytes),
     "ar_create_all");
  std::vector<std::unique_ptr<ProcessGroup>> groups;
  for (int r = 0; r < world; ++r)
    groups.emplace_back(std::make_unique<ProcessGroupB200>(r, world, devices[r], comms[r]));
  return groups;
}

ProcessGroupB200::~ProcessGroupB200() { b200_ar_destroy(comm_); }

void ProcessGroupB200::allreduce(torch::Tensor& input) const {
  TORCH_CHECK(input.device() == device(), "input should be on the same device as the process group");
  if (world_size() == 1 || input.numel() == 0) return;
  c10::cuda::CUDAGuard guard(device());
  // Any size, like the reference's NCCL group (process_group.cpp:135-153): messages larger than
  // the symmetric buffer go through it in slices; a tensor that is not contiguous, not 16-byte
  // aligned or not a multiple of 16 bytes is staged through a zero-padded temporary.  Every
  // decision below depends on size / dtype only, so all ranks take the same path.
  const int64_t es = static_cast<int64_t>(input.element_size());
  const int64_t bytes = input.numel() * es;
  const bool direct = input.is_contiguous() && bytes % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(input.data_ptr()) % 16 == 0;
  torch::Tensor staged;
  if (!direct) {
    const int64_t padded = (bytes + 15) / 16 * 16 / es;
    staged = torch::zeros({padded}, input.options());
    staged.narrow(0, 0, input.numel()).copy_(input.reshape({-1}));
  }
  auto* base = static_cast<uint8_t*>(direct ? input.data_ptr() : staged.data_ptr());
  const int64_t total = direct ? input.numel() : staged.numel();
  const int64_t slice = (kMaxBytes / 2) / es;  // elements per call (a multiple of 16 bytes)
  const auto stream = at::cuda::getCurrentCUDAStream().stream();
  for (int64_t off = 0; off < total; off += slice) {
    const int64_t n = std::min(slice, total - off);
    ok(b200_ar_allreduce(comm_, base + off * es, n, dtype_code(input), stream), "ar_allreduce");
  }
  if (!direct) input.copy_(staged.narrow(0, 0, input.numel()).view_as(input));
}

torch::Tensor ProcessGroupB200::allgather_lastdim(const torch::Tensor& input) const {
  TORCH_CHECK(input.device() == device(), "input should be on the same device as the process group");
  const auto x = input.contiguous();
  auto sizes = x.sizes().vec();
  TORCH_CHECK(!sizes.empty(), "allgather_lastdim: needs at least one dimension");
  const int64_t cols = sizes.back();
  sizes.back() = cols * world_size();
  auto out = torch::empty(sizes, x.options());
  if (x.numel() == 0) return out;
  c10::cuda::CUDAGuard guard(device());
  const int64_t row_bytes = cols * static_cast<int64_t>(x.element_size());
  const int64_t rows = x.numel() / cols;
  TORCH_CHECK(row_bytes % 16 == 0 && row_bytes <= kMaxBytes,
              "allgather_lastdim: rows of a multiple of 16 bytes, at most ", kMaxBytes, " bytes");
  // more rows than the symmetric buffer holds: gather them in groups
  const int64_t rows_per_call = std::max<int64_t>(1, kMaxBytes / row_bytes);
  const auto stream = at::cuda::getCurrentCUDAStream().stream();
  for (int64_t r0 = 0; r0 < rows; r0 += rows_per_call) {
    const int64_t nr = std::min(rows_per_call, rows - r0);
    ok(b200_ar_allgather(comm_, static_cast<uint8_t*>(out.data_ptr()) + r0 * row_bytes * world_size(),
                         static_cast<const uint8_t*>(x.const_data_ptr()) + r0 * row_bytes, nr, row_bytes,
                         stream),
       "ar_allgather");
  }
  return out;
}

torch::Tensor ProcessGroupB200::allreduce_partials(const torch::Tensor& partials, int64_t gemm_k,
                                                   torch::ScalarType dtype) const {
  TORCH_CHECK(partials.dim() == 3 && partials.scalar_type() == torch::kFloat && partials.is_contiguous(),
              "allreduce_partials: fp32 [slots, rows, n]");
  const int64_t slots = partials.size(0), rows = partials.size(1), n = partials.size(2);
  auto out = torch::empty({rows, n}, partials.options().dtype(dtype));
  if (rows == 0) return out;
  c10::cuda::CUDAGuard guard(device());
  ok(b200_ar_allreduce_splitk(comm_, out.data_ptr(), partials.const_data_ptr<float>(),
                              static_cast<int>(slots), gemm_k, n, rows * n, dtype_code(out),
                              at::cuda::getCurrentCUDAStream().stream()),
     "ar_allreduce_splitk");
  return out;
}

bool ProcessGroupB200::supports_partials_norm(int64_t rows, int64_t n, torch::ScalarType dtype) const {
  // two-shot form (default; LL lines: twice the payload in the buffers): rows <= 128, n <= 8192;
  // one-shot (B200_AR_ALGO=oneshot): rows <= 64, n <= 4096
  const char* algo = std::getenv("B200_AR_ALGO");
  const bool two = !(algo && std::string(algo) == "oneshot");
  return world_size() > 1 && rows > 0 && rows <= (two ? 128 : 64) && n % 128 == 0 &&
         n <= (two ? 8192 : 4096) && (rows + world_size()) * n * 2 * (two ? 2 : 1) <= kMaxBytes &&
         (dtype == torch::kBFloat16 || dtype == torch::kHalf);
}

torch::Tensor ProcessGroupB200::allreduce_partials_norm(const torch::Tensor& partials, int64_t gemm_k,
                                                        torch::Tensor& residual,
                                                        const torch::Tensor& weight, float eps) const {
  TORCH_CHECK(partials.dim() == 3 && partials.scalar_type() == torch::kFloat && partials.is_contiguous() &&
                  residual.is_contiguous(),
              "allreduce_partials_norm: fp32 [slots, rows, n] partials, contiguous residual");
  const int64_t slots = partials.size(0), rows = partials.size(1), n = partials.size(2);
  auto out = torch::empty_like(residual);
  c10::cuda::CUDAGuard guard(device());
  ok(b200_ar_allreduce_splitk_norm(comm_, out.data_ptr(), residual.data_ptr(),
                                   partials.const_data_ptr<float>(), static_cast<int>(slots), gemm_k,
                                   weight.const_data_ptr(), rows, n, eps, dtype_code(residual),
                                   at::cuda::getCurrentCUDAStream().stream()),
     "ar_allreduce_splitk_norm");
  return out;
}

void ProcessGroupB200::allgather(const torch::Tensor& input, torch::Tensor& outputs) const {
  // cat along dim 0 == "last dim" gather of the flattened tensor seen as one row
  TORCH_CHECK(outputs.is_contiguous() && outputs.numel() == input.numel() * world_size() &&
                  outputs.scalar_type() == input.scalar_type(),
              "allgather: outputs must be contiguous with world_size * input.numel() elements");
  const auto x = input.contiguous();
  if (x.numel() == 0) return;
  c10::cuda::CUDAGuard guard(device());
  const int64_t bytes = x.numel() * static_cast<int64_t>(x.element_size());
  TORCH_CHECK(bytes % 16 == 0, "allgather: message must be a multiple of 16 bytes");
  if (bytes <= kMaxBytes) {
    ok(b200_ar_allgather(comm_, outputs.data_ptr(), x.const_data_ptr(), 1, bytes,
                         at::cuda::getCurrentCUDAStream().stream()),
       "ar_allgather");
    return;
  }
  // larger than the symmetric buffer: slices land at [r][off .. off+n) of the [world, bytes] result
  auto tmp = torch::empty({world_size() * kMaxBytes}, x.options().dtype(torch::kUInt8));
  auto out_b = outputs.view({-1}).view(torch::kUInt8).view({world_size(), bytes});
  const auto* src = static_cast<const uint8_t*>(x.const_data_ptr());
  for (int64_t off = 0; off < bytes; off += kMaxBytes) {
    const int64_t n = std::min<int64_t>(kMaxBytes, bytes - off);
    ok(b200_ar_allgather(comm_, tmp.data_ptr(), src + off, 1, n, at::cuda::getCurrentCUDAStream().stream()),
       "ar_allgather");
    out_b.narrow(1, off, n).copy_(tmp.narrow(0, 0, world_size() * n).view({world_size(), n}));
  }
}

void ProcessGroupB200::allgather(const torch::Tensor& input,
                                 std::vector<torch::Tensor>& outputs) const {
  TORCH_CHECK(static_cast<int>(outputs.size()) == world_size(),
              "outputs should have the same size as world_size");
  auto flat = torch::empty({world_size(), input.numel()}, input.options());
  allgather(input, flat);
  for (int r = 0; r < world_size(); ++r) outputs[r].copy_(flat[r].view_as(input));
}

torch::Tensor gather_from_model_parallel_region(const torch::Tensor& input, const ParallelArgs& pa) {
  if (pa.world_size() == 1) return input;
  auto* pg = dynamic_cast<ProcessGroupB200*>(pa.process_group());
  if (pg != nullptr) return pg->allgather_lastdim(input);
  std::vector<torch::Tensor> parts;
  for (int r = 0; r < pa.world_size(); ++r) parts.push_back(torch::empty_like(input));
  pa.process_group()->allgather(input, parts);
  return torch::cat(parts, /*dim=*/-1).contiguous();
}

torch::Tensor reduce_from_model_parallel_region(torch::Tensor input, const ParallelArgs& pa) {
  if (pa.world_size() == 1) return input;
  pa.process_group()->allreduce(input);
  return input;
}

torch::Tensor scatter_to_model_parallel_region(const torch::Tensor& input, const ParallelArgs& pa) {
  if (pa.world_size() == 1) return input;
  const int64_t last = input.size(-1);
  TORCH_CHECK(last % pa.world_size() == 0, "last_dim_size ", last, " not divisible by world_size ",
              pa.world_size());
  return input.split(last / pa.world_size(), /*dim=*/-1)[pa.rank()];
}

}  // namespace llm
