// decode_demo.cpp — the decode step driven from C++ only (no Python in the process): random-init
// Llama-3-8B-shaped int4 model (any layer count), a paged KV cache with shuffled block ids, and the
// ModelRunner's captured graph replayed for `steps` decode steps.  Prints one line:
//   decode_demo: layers=L batch=B kv_len=S steps=K  ms/step=T  tokens/s=V  (replayed=R eager=E)
// It shows what a C++ engine links against (shim/b200_layers.h + libb200decode.so); the measured
// benchmark of the repo is bench.py.
//   build: __graft_entry__._build_shim() also links scalellm_b200/decode_demo
//   run:   scalellm_b200/decode_demo [layers=4] [batch=64] [kv_len=2048] [steps=20] [tp=1]
// tp > 1: the reference engine's threading model — one process, ProcessGroup::create_process_groups
// over the first tp devices, one worker thread and one LlamaDecoderStep (its shard) per GPU.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/torch.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#include "b200_layers.h"

namespace {

torch::Tensor random_words(std::vector<int64_t> shape, const torch::Device& dev) {
  // any 32-bit word is eight valid int4 weights
  return torch::randint(-2147483647LL - 1, 2147483647LL, shape,
                        torch::dtype(torch::kInt64).device(dev))
      .to(torch::kInt32);
}

void add_awq_linear(llm::StateDict& sd, const std::string& prefix, int64_t K, int64_t N,
                    const torch::Device& dev) {
  const int64_t g = 128;
  sd[prefix + "qweight"] = random_words({K, N / 8}, dev);
  sd[prefix + "qzeros"] = random_words({K / g, N / 8}, dev);
  sd[prefix + "scales"] =
      (torch::randn({K / g, N}, torch::dtype(torch::kFloat).device(dev)).abs() * 0.01 + 1e-4).to(torch::kBFloat16);
}

struct Config {
  int64_t n_layers, B, S, steps;
  int world;
};

// host barrier between the worker threads (device work is ordered by the collectives themselves)
struct SpinBarrier {
  explicit SpinBarrier(int n) : n_(n) {}
  void wait() {
    const int gen = gen_.load();
    if (count_.fetch_add(1) + 1 == n_) {
      count_.store(0);
      gen_.fetch_add(1);
    } else {
      while (gen_.load() == gen) std::this_thread::yield();
    }
  }
  int n_;
  std::atomic<int> count_{0}, gen_{0};
};

void worker(int rank, const Config& cfg, const llm::LlamaArgs& args, const llm::StateDict& sd,
            const torch::Tensor& inv_freq, llm::ProcessGroup* pg, const torch::Tensor& table_dev0,
            const torch::Tensor& tokens_dev0, SpinBarrier* barrier) {
  torch::NoGradGuard no_grad;
  const torch::Device dev(torch::kCUDA, static_cast<c10::DeviceIndex>(rank));
  c10::cuda::CUDAGuard guard(dev);
  const auto bf16 = torch::dtype(torch::kBFloat16).device(dev);
  const auto i32 = torch::dtype(torch::kInt32).device(dev);
  const int64_t B = cfg.B, S = cfg.S, steps = cfg.steps, D = args.head_dim;
  llm::QuantArgs qa;
  qa.quant_method = "awq";
  llm::LlamaDecoderStep model(args, qa, inv_freq, bf16, llm::ParallelArgs(rank, cfg.world, pg));
  model.load_state_dict(sd);  // keeps this rank's shard (copied to its device)
  // One process, one thread per GPU: every cudaMalloc synchronises with the peer-mapped devices, so
  // none may happen while another rank's all-reduce kernel spins for this rank.  Repack the weights
  // now and leave the caching allocator a block large enough for the step's temporaries.
  model.prepack();
  { auto reserve = torch::empty({int64_t(1) << 30}, torch::dtype(torch::kByte).device(dev)); }

  // paged KV cache: block_size 8, every sequence at kv_len S with room to grow; this rank's kv heads
  const int64_t bs = 8, cap = S + steps + 8, blocks_per_seq = (cap + bs - 1) / bs;
  const int64_t n_blocks = B * blocks_per_seq + 1;  // block 0 is left to the graph capture's warm-up
  const int64_t kv_heads = std::max<int64_t>(1, args.n_kv_heads / cfg.world);
  std::vector<llm::KVCache> caches;
  for (int64_t i = 0; i < cfg.n_layers; ++i)
    caches.emplace_back(torch::randn({n_blocks * bs, kv_heads, D}, bf16),
                        torch::randn({n_blocks * bs, kv_heads, D}, bf16), bs);
  model.set_kv_caches(std::move(caches));

  llm::ModelRunner::Options ro;
  ro.cuda_graph_batch_sizes = {static_cast<uint32_t>(B)};
  ro.cuda_graph_max_seq_len = cap;
  ro.block_size = bs;
  ro.greedy = true;
  llm::ModelRunner runner(&model, dev, ro);
  barrier->wait();  // every rank's model is built: the capture's warm-up step runs collectives
  runner.capture_cuda_graphs(static_cast<uint32_t>(B));

  const torch::Tensor table_full = table_dev0.to(dev);  // same block ids on every rank
  torch::Tensor tokens = tokens_dev0.to(dev);
  auto step = [&](int64_t kv_len) {  // every sequence decodes one token at position kv_len - 1
    llm::InputParameters p;
    p.num_sequences = static_cast<int32_t>(B);
    p.q_max_seq_len = 1;
    p.kv_max_seq_len = static_cast<int32_t>(kv_len);
    p.q_cu_seq_lens = torch::arange(0, B + 1, i32);
    p.kv_cu_seq_lens = torch::arange(0, B + 1, i32) * static_cast<int32_t>(kv_len);
    const int64_t nb = (kv_len + bs - 1) / bs;
    p.block_tables = table_full.slice(1, 0, nb).contiguous().view({-1});
    p.cu_block_lens = torch::arange(0, B + 1, i32) * static_cast<int32_t>(nb);
    p.new_cache_slots = table_full.select(1, (kv_len - 1) / bs) + static_cast<int32_t>((kv_len - 1) % bs);
    const torch::Tensor positions = torch::full({B}, kv_len - 1, i32);
    tokens = runner.forward(tokens, positions, p).to(torch::kInt32);  // greedy: next ids feed the next step
  };
  int64_t kv = S;
  for (int i = 0; i < 3; ++i) step(++kv);
  torch::cuda::synchronize();
  barrier->wait();
  const auto stream = at::cuda::getCurrentCUDAStream();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  for (int64_t i = 0; i < steps; ++i) step(++kv);
  cudaEventRecord(e1, stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  barrier->wait();
  if (rank == 0)
    std::printf("decode_demo: layers=%lld batch=%lld kv_len=%lld steps=%lld tp=%d  ms/step=%.3f  tokens/s=%.0f  "
                "(replayed=%lld eager=%lld)\n",
                (long long)cfg.n_layers, (long long)B, (long long)S, (long long)steps, cfg.world, ms / steps,
                1000.0 * B * steps / ms, (long long)runner.num_cuda_graph_replayed(),
                (long long)runner.num_eager_execution());
}

}  // namespace

int main(int argc, char** argv) {
  Config cfg;
  cfg.n_layers = argc > 1 ? std::atoll(argv[1]) : 4;
  cfg.B = argc > 2 ? std::atoll(argv[2]) : 64;
  cfg.S = argc > 3 ? std::atoll(argv[3]) : 2048;
  cfg.steps = argc > 4 ? std::atoll(argv[4]) : 20;
  cfg.world = argc > 5 ? std::atoi(argv[5]) : 1;
  if (!torch::cuda::is_available()) {
    std::fprintf(stderr, "decode_demo: no CUDA device (there is no CPU path)\n");
    return 2;
  }
  if (cfg.world < 1 || cfg.world > static_cast<int>(torch::cuda::device_count())) {
    std::fprintf(stderr, "decode_demo: tp=%d needs that many CUDA devices\n", cfg.world);
    return 2;
  }
  torch::NoGradGuard no_grad;
  const torch::Device dev(torch::kCUDA, 0);
  c10::cuda::CUDAGuard guard(dev);
  torch::manual_seed(0);

  llm::LlamaArgs args;  // Llama-3-8B shapes
  args.n_layers = cfg.n_layers;
  const int64_t h = args.hidden_size, D = args.head_dim, I = args.intermediate_size;
  const int64_t qkv_n = (args.n_heads + 2 * args.n_kv_heads) * D;
  llm::StateDict sd;  // the unsharded checkpoint, on device 0; every rank cuts and copies its shard
  for (int64_t i = 0; i < cfg.n_layers; ++i) {
    const std::string p = "layers." + std::to_string(i) + ".";
    add_awq_linear(sd, p + "qkv.", h, qkv_n, dev);
    add_awq_linear(sd, p + "o.", args.n_heads * D, h, dev);
    add_awq_linear(sd, p + "gate_up.", h, 2 * I, dev);
    add_awq_linear(sd, p + "down.", I, h, dev);
    sd[p + "input_norm.weight"] = torch::ones({h}, torch::dtype(torch::kBFloat16).device(dev));
    sd[p + "post_norm.weight"] = torch::ones({h}, torch::dtype(torch::kBFloat16).device(dev));
  }
  const auto bf16 = torch::dtype(torch::kBFloat16).device(dev);
  sd["final_norm.weight"] = torch::ones({h}, bf16);
  sd["embed.weight"] = (torch::randn({args.vocab_size, h}, torch::dtype(torch::kFloat).device(dev)) * 0.02).to(torch::kBFloat16);
  sd["lm_head.weight"] = (torch::randn({args.vocab_size, h}, torch::dtype(torch::kFloat).device(dev)) * 0.02).to(torch::kBFloat16);
  // inverse frequencies of the rotary embedding (theta 5e5, no scaling: timing only)
  const torch::Tensor inv_freq =
      1.0 / torch::pow(500000.0, torch::arange(0, D, 2, torch::kFloat) / static_cast<double>(D));

  // first-slot ids of a random permutation of blocks 1..n_blocks-1, blocks_per_seq per sequence
  const int64_t bs = 8, cap = cfg.S + cfg.steps + 8, blocks_per_seq = (cap + bs - 1) / bs;
  const int64_t n_blocks = cfg.B * blocks_per_seq + 1;
  const torch::Tensor table =
      ((torch::randperm(n_blocks - 1, torch::dtype(torch::kInt64).device(dev)) + 1) * bs).to(torch::kInt32)
          .slice(0, 0, cfg.B * blocks_per_seq).view({cfg.B, blocks_per_seq});
  const torch::Tensor tokens = torch::randint(0, args.vocab_size, {cfg.B}, torch::dtype(torch::kInt32).device(dev));
  torch::cuda::synchronize();

  std::vector<std::unique_ptr<llm::ProcessGroup>> groups;
  if (cfg.world > 1) {
    std::vector<torch::Device> devices;
    for (int r = 0; r < cfg.world; ++r) devices.emplace_back(torch::kCUDA, static_cast<c10::DeviceIndex>(r));
    groups = llm::ProcessGroup::create_process_groups(devices);
  }
  SpinBarrier barrier(cfg.world);
  std::vector<std::thread> threads;
  for (int r = 0; r < cfg.world; ++r)
    threads.emplace_back(worker, r, std::cref(cfg), std::cref(args), std::cref(sd), std::cref(inv_freq),
                         cfg.world > 1 ? groups[r].get() : nullptr, std::cref(table), std::cref(tokens), &barrier);
  for (auto& t : threads) t.join();
  return 0;
}
