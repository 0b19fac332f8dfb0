// b200_process_group.h — the reference's ProcessGroup interface and model-parallel helpers over the
// NVLink peer-memory collectives of libb200decode, in the reference engine's threading model: ONE
// process, one communicator (and one worker thread) per GPU.
//
//   ProcessGroup (+ ProcessGroupB200)        src/model_parallel/process_group.h:10-60,
//                                            process_group.cpp:98-153 (create_process_groups =
//                                            ncclCommInitAll, allreduce, allgather)
//   ParallelArgs                             src/model_parallel/parallel_args.h:10-22
//   gather_from / reduce_from / scatter_to_model_parallel_region
//                                            src/model_parallel/model_parallel.cpp:13-65
//
// Inside ScaleLLM ProcessGroupB200 derives from ProcessGroupNCCL and keeps NCCL for what the
// decode step does not use (alltoall).  Messages of any size and layout are
// taken: larger than the symmetric buffer -> slices, unaligned / non-contiguous -> staged.
#pragma once

#include <torch/torch.h>

#include <memory>
#include <vector>

struct b200_ar_comm;

namespace llm {

class ProcessGroup {
 public:
  ProcessGroup(int rank, int world_size, const torch::Device& device)
      : rank_(rank), world_size_(world_size), device_(device) {}
  virtual ~ProcessGroup() = default;

  int rank() const { return rank_; }
  int world_size() const { return world_size_; }
  const torch::Device& device() const { return device_; }

  // in-place sum over the group, on the current stream of this group's device
  virtual void allreduce(torch::Tensor& input) const = 0;
  // outputs[r] = rank r's input
  virtual void allgather(const torch::Tensor& input, std::vector<torch::Tensor>& outputs) const = 0;
  // outputs = cat over ranks along dim 0
  virtual void allgather(const torch::Tensor& input, torch::Tensor& outputs) const = 0;

  // one group per device, all in this process (process_group.cpp:98-118)
  static std::vector<std::unique_ptr<ProcessGroup>> create_process_groups(
      const std::vector<torch::Device>& devices);

 private:
  int rank_ = 0;
  int world_size_ = 0;
  torch::Device device_;
};

class ProcessGroupB200 final : public ProcessGroup {
 public:
  // bytes of the largest single launch of the peer-memory kernels (the symmetric region is four
  // times that); larger messages are sliced
  static constexpr int64_t kMaxBytes = 16 << 20;

  ProcessGroupB200(int rank, int world_size, const torch::Device& device, b200_ar_comm* comm)
      : ProcessGroup(rank, world_size, device), comm_(comm) {}
  ~ProcessGroupB200() override;

  void allreduce(torch::Tensor& input) const override;
  void allgather(const torch::Tensor& input, std::vector<torch::Tensor>& outputs) const override;
  void allgather(const torch::Tensor& input, torch::Tensor& outputs) const override;

  // cat(all-gather(input), dim=-1) in one launch (what gather_from_model_parallel_region needs)
  torch::Tensor allgather_lastdim(const torch::Tensor& input) const;

  // Row-parallel W4A16 GEMM -> all-reduce without the GEMM's own reduction pass: `partials` are
  // this rank's fp32 stream-K partials [slots, rows, n] of a [gemm_k, n] weight
  // (include/b200_decode.h "partials mode"); returns the reduced [rows, n] in `dtype`.
  torch::Tensor allreduce_partials(const torch::Tensor& partials, int64_t gemm_k,
                                   torch::ScalarType dtype) const;
  // ... with the residual add and the RMSNorm that follow fused in as well (one launch):
  // residual += all-reduced row; returns rms_norm(residual) * weight
  bool supports_partials_norm(int64_t rows, int64_t n, torch::ScalarType dtype) const;
  torch::Tensor allreduce_partials_norm(const torch::Tensor& partials, int64_t gemm_k,
                                        torch::Tensor& residual, const torch::Tensor& weight,
                                        float eps) const;

  b200_ar_comm* comm() const { return comm_; }

 private:
  b200_ar_comm* comm_ = nullptr;
};

class ParallelArgs {  // parallel_args.h:10-22: the same three accessors
 public:
  ParallelArgs(int32_t rank, int32_t world_size, ProcessGroup* process_group)
      : rank_(rank), world_size_(world_size), process_group_(process_group) {}
  int32_t rank() const { return rank_; }
  int32_t world_size() const { return world_size_; }
  ProcessGroup* process_group() const { return process_group_; }  // nullptr if world size is 1

 private:
  int32_t rank_ = 0;
  int32_t world_size_ = 1;
  ProcessGroup* process_group_ = nullptr;
};

torch::Tensor gather_from_model_parallel_region(const torch::Tensor& input, const ParallelArgs& pa);
torch::Tensor reduce_from_model_parallel_region(torch::Tensor input, const ParallelArgs& pa);
torch::Tensor scatter_to_model_parallel_region(const torch::Tensor& input, const ParallelArgs& pa);

}  // namespace llm
