// allreduce.cu — one-shot tensor-parallel all-reduce over NVLink peer memory.
//
// Replaces ProcessGroupNCCL::allreduce (src/model_parallel/process_group.cpp:135-153)
// for the latency-bound <= 1 MiB row-parallel reductions of the decode step
// (2 per layer: after o_proj and down_proj).
//
// One process per GPU (b200_ar_create + b200_ar_open_peers, CUDA IPC), or all ranks in one
// process with one thread per GPU like the reference's engine (b200_ar_create_all, the
// counterpart of its ncclCommInitAll, process_group.cpp:98-118: peer access instead of IPC).
// Each rank owns ONE cudaMalloc'ed symmetric region
//   [ data buffer, parity 0 | data buffer, parity 1 | flags[world][AR_MAX_BLOCKS] | epoch | done ]
// exported through CUDA IPC and mapped by every peer (NVSwitch gives every pair
// full NVLink bandwidth).  A call with epoch e:
//   1. every block copies its slice of the input into the local buffer (e & 1),
//   2. publishes flag value e into every peer's flag array (st.release.sys),
//   3. waits until all peers' flags for the same block index reach e,
//   4. sums the slice over ranks in rank order 0..w-1 in fp32 (identical order on
//      every rank => bit-identical results on all ranks) and writes it in place.
// Double buffering by epoch parity plus the flag barrier of the next call makes
// reuse safe: nobody can overwrite buffer (e & 1) before every peer finished
// reading it in call e (they must have entered call e+1 first).  The epoch lives
// in device memory and is advanced by the kernel, so a captured CUDA graph can
// be replayed (no host-side state baked into kernel arguments).
//
// Fused forms (the row-parallel GEMM -> all-reduce -> residual add -> RMSNorm chain of a TP
// decoder layer, models/meta/llama.h:170-177):
//   * b200_ar_allreduce_splitk: step 1 reads the producing W4A16 GEMM's fp32 stream-K partials
//     and sums each tile's contributor slots (common.cuh W4Plan) instead of copying a bf16 input;
//   * b200_ar_allreduce_splitk_norm: additionally, one block per row, step 4 feeds the reduced row
//     straight into residual += x; out = rms_norm(residual) * w — one launch for the whole chain,
//     bit-identical to the separate kernels (same rounding points).

#include <cstring>

#include "common.cuh"

namespace b200 {
constexpr int AR_MAX_WORLD = 8;
constexpr int AR_MAX_BLOCKS = 64;
constexpr int AR_THREADS = 512;
}  // namespace b200

struct b200_ar_comm {
  int rank = 0, world = 1;
  int device = 0;
  int64_t max_bytes = 0;
  uint8_t* local = nullptr;                 // base of the local symmetric region
  uint8_t* peer[b200::AR_MAX_WORLD] = {};   // mapped bases, peer[rank] == local
  bool opened = false;
  bool ipc = true;   // peers mapped through CUDA IPC (one process per GPU); false: same process
};

namespace b200 {

struct ArDevPtrs {
  uint8_t* base[AR_MAX_WORLD];
};

__host__ __device__ inline int64_t ar_flags_off(int64_t max_bytes) { return 2 * max_bytes; }
__host__ __device__ inline int64_t ar_epoch_off(int64_t max_bytes) {
  return 2 * max_bytes + (int64_t)AR_MAX_WORLD * AR_MAX_BLOCKS * 4;
}
__host__ __device__ inline int64_t ar_region_bytes(int64_t max_bytes) {
  return ar_epoch_off(max_bytes) + 256;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

template <typename T>
__device__ __forceinline__ void accum16(float (&acc)[16 / sizeof(T)], uint4 v) {
  const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(16 / sizeof(T)); ++i) acc[i] += Num<T>::to_f(e[i]);
}

// partials != nullptr: the local input is the producing GEMM's stream-K partials [slots][count]
// fp32 (row length row_n); the copy-in stage sums each tile's contributor slots (fixed order) and
// rounds once to T — the GEMM epilogue's job.
// NORM: grid = one block per row (row_n / VEC <= AR_THREADS vectors); the reduced row is not stored
// but consumed in place: residual += T(sum over ranks); out = rms_norm(residual) * weight — the
// o_proj / down_proj all-reduce, the residual add and the following RMSNorm in ONE launch,
// bit-identical to the three separate kernels (same rounding points).
template <typename T>
struct ArNormArgs {
  T* residual;
  const T* weight;
  T* out;
  float eps;
};

template <typename T, bool NORM>
__global__ void __launch_bounds__(AR_THREADS) allreduce_oneshot_kernel(ArDevPtrs ptrs, T* data,
                                                                      int64_t nvec, int rank,
                                                                      int world,
                                                                      int64_t max_bytes,
                                                                      const float* partials,
                                                                      W4Plan plan, int row_n,
                                                                      int64_t split_stride,
                                                                      ArNormArgs<T> na) {
  constexpr int VEC = 16 / sizeof(T);
  uint8_t* local = ptrs.base[rank];
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(local + ar_epoch_off(max_bytes));
  uint32_t* done_ptr = epoch_ptr + 1;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int64_t buf_off = (e & 1) ? max_bytes : 0;

  // slice of this block (in 16-byte vectors)
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = (int64_t)blockIdx.x * per;
  const int64_t v1 = v0 + per < nvec ? v0 + per : nvec;

  // 1. stage my slice
  uint4* mine = reinterpret_cast<uint4*>(local + buf_off);
  const uint4* src = reinterpret_cast<const uint4*>(data);
  if (partials == nullptr) {
    for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) mine[i] = src[i];
  } else if constexpr (sizeof(T) == 2) {
    for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) {
      float a[8];
      const int col = (int)((i * 8) % row_n);  // 8 consecutive columns of one row, inside one n tile
      w4_sum_partials8(a, partials + i * 8, split_stride, w4_contrib_col(plan, col));
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) oe[k] = Num<T>::from_f(a[k]);
      mine[i] = o;
    }
  }
  __threadfence_system();
  __syncthreads();

  // 2. tell every peer this slice of mine is ready; 3. wait for theirs
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    uint32_t* their_flags = reinterpret_cast<uint32_t*>(ptrs.base[r] + ar_flags_off(max_bytes));
    st_release_sys(&their_flags[rank * AR_MAX_BLOCKS + blockIdx.x], e);
    const uint32_t* my_flags =
        reinterpret_cast<const uint32_t*>(local + ar_flags_off(max_bytes));
    while ((int32_t)(ld_acquire_sys(&my_flags[r * AR_MAX_BLOCKS + blockIdx.x]) - e) < 0) {
    }
  }
  __syncthreads();

  // 4. reduce in rank order, write back in place (or feed the residual + RMSNorm epilogue)
  if constexpr (!NORM) {
    for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) {
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      for (int r = 0; r < world; ++r) {
        const uint4* pb = reinterpret_cast<const uint4*>(ptrs.base[r] + buf_off);
        accum16<T>(acc, ld_volatile_v4(pb + i));
      }
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int k = 0; k < VEC; ++k) oe[k] = Num<T>::from_f(acc[k]);
      reinterpret_cast<uint4*>(data)[i] = o;
    }
  } else {
    __shared__ float red[32];
    const int64_t i = v0 + threadIdx.x;  // one vector per thread: the block's slice is one row
    const bool have = i < v1;
    float x[VEC];
    float ss = 0.f;
    if (have) {
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      for (int r = 0; r < world; ++r) {
        const uint4* pb = reinterpret_cast<const uint4*>(ptrs.base[r] + buf_off);
        accum16<T>(acc, ld_volatile_v4(pb + i));
      }
      uint4 rraw = reinterpret_cast<const uint4*>(na.residual)[i];
      const T* rr = reinterpret_cast<const T*>(&rraw);
      uint4 sraw;
      T* sv = reinterpret_cast<T*>(&sraw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float reduced = rnd<T>(acc[k]);  // what the plain all-reduce would have stored
        const float f = Num<T>::to_f(rr[k]) + reduced;
        ss += f * f;
        sv[k] = Num<T>::from_f(f);
        x[k] = Num<T>::to_f(sv[k]);
      }
      reinterpret_cast<uint4*>(na.residual)[i] = sraw;
    }
    const float total = block_sum<AR_THREADS>(ss, red);
    const float rstd = rsqrtf(total / row_n + na.eps);
    if (have) {
      const int col = (int)((i * VEC) % row_n);
      uint4 wraw = *reinterpret_cast<const uint4*>(na.weight + col);
      const T* w = reinterpret_cast<const T*>(&wraw);
      uint4 oraw;
      T* o = reinterpret_cast<T*>(&oraw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float y = rnd<T>(x[k] * rstd);
        o[k] = Num<T>::from_f(y * Num<T>::to_f(w[k]));
      }
      reinterpret_cast<uint4*>(na.out)[i] = oraw;
    }
  }

  // advance the epoch once every block is done with it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t d = atomicAdd(done_ptr, 1u);
    if (d == gridDim.x - 1) {
      *done_ptr = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(epoch_ptr) = e;
    }
  }
}

// All-gather with the same staging / flag protocol (and the same epoch counter, so all-reduces and
// all-gathers of one communicator may be mixed freely as long as every rank issues the same
// sequence): rank r's input [rows][cols_v vectors] lands in out[row][r * cols_v + c] on every
// rank — the layout gather_from_model_parallel_region builds with allgather + cat(dim=-1)
// (model_parallel.cpp:13-31), without the temporaries.  Pure 16-byte copies: bit exact.
__global__ void __launch_bounds__(AR_THREADS) allgather_oneshot_kernel(ArDevPtrs ptrs,
                                                                      const uint4* __restrict__ in,
                                                                      uint4* __restrict__ out,
                                                                      int64_t nvec, int64_t cols_v,
                                                                      int rank, int world,
                                                                      int64_t max_bytes) {
  uint8_t* local = ptrs.base[rank];
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(local + ar_epoch_off(max_bytes));
  uint32_t* done_ptr = epoch_ptr + 1;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int64_t buf_off = (e & 1) ? max_bytes : 0;
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = (int64_t)blockIdx.x * per;
  const int64_t v1 = v0 + per < nvec ? v0 + per : nvec;

  uint4* mine = reinterpret_cast<uint4*>(local + buf_off);
  for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) mine[i] = in[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    uint32_t* their_flags = reinterpret_cast<uint32_t*>(ptrs.base[r] + ar_flags_off(max_bytes));
    st_release_sys(&their_flags[rank * AR_MAX_BLOCKS + blockIdx.x], e);
    const uint32_t* my_flags = reinterpret_cast<const uint32_t*>(local + ar_flags_off(max_bytes));
    while ((int32_t)(ld_acquire_sys(&my_flags[r * AR_MAX_BLOCKS + blockIdx.x]) - e) < 0) {
    }
  }
  __syncthreads();
  for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) {
    const int64_t row = i / cols_v, c = i - row * cols_v;
    for (int r = 0; r < world; ++r) {
      const uint4* pb = reinterpret_cast<const uint4*>(ptrs.base[r] + buf_off);
      out[(row * world + r) * cols_v + c] = ld_volatile_v4(pb + i);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t d = atomicAdd(done_ptr, 1u);
    if (d == gridDim.x - 1) {
      *done_ptr = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(epoch_ptr) = e;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_ar_allgather(b200_ar_comm* c, void* out, const void* in, int64_t rows, int64_t row_bytes,
                      b200_stream_t stream) {
  B200_CHECK_ARG(c && out && in, "ar_allgather: null pointer");
  B200_CHECK_ARG(rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0 && is_aligned(out, 16) &&
                     is_aligned(in, 16),
                 "ar_allgather: rows of a multiple of 16 bytes, 16-byte aligned buffers");
  if (rows == 0) return B200_OK;
  auto st = static_cast<cudaStream_t>(stream);
  if (c->world == 1) {
    if (out != in)
      B200_CUDA_OK(cudaMemcpyAsync(out, in, (size_t)(rows * row_bytes), cudaMemcpyDeviceToDevice, st));
    return B200_OK;
  }
  B200_CHECK_ARG(c->opened, "ar_allgather: peers not opened");
  B200_CHECK_ARG(out != in, "ar_allgather: out must not alias in");
  const int64_t bytes = rows * row_bytes;
  if (bytes > c->max_bytes)
    return set_error(B200_ERR_WORKSPACE, "ar_allgather: %lld B exceeds the %lld B symmetric buffer",
                     (long long)bytes, (long long)c->max_bytes);
  const int64_t nvec = bytes / 16;
  int blocks = (int)((nvec + 2 * AR_THREADS - 1) / (2 * AR_THREADS));
  if (blocks < 1) blocks = 1;
  if (blocks > AR_MAX_BLOCKS) blocks = AR_MAX_BLOCKS;
  ArDevPtrs ptrs{};
  for (int r = 0; r < c->world; ++r) ptrs.base[r] = c->peer[r];
  allgather_oneshot_kernel<<<blocks, AR_THREADS, 0, st>>>(
      ptrs, static_cast<const uint4*>(in), static_cast<uint4*>(out), nvec, row_bytes / 16, c->rank,
      c->world, c->max_bytes);
  B200_LAUNCH_OK("allgather_oneshot");
  return B200_OK;
}

int b200_ar_create(b200_ar_comm** comm, int rank, int world_size, int64_t max_bytes,
                   void* handle_out) {
  B200_CHECK_ARG(comm && handle_out, "ar_create: null pointer");
  B200_CHECK_ARG(world_size >= 1 && world_size <= AR_MAX_WORLD && rank >= 0 && rank < world_size,
                 "ar_create: rank %d / world %d unsupported (max %d)", rank, world_size,
                 AR_MAX_WORLD);
  B200_CHECK_ARG(max_bytes > 0 && max_bytes % 16 == 0, "ar_create: max_bytes must be a multiple of 16");
  auto* c = new b200_ar_comm();
  c->rank = rank;
  c->world = world_size;
  c->max_bytes = max_bytes;
  B200_CUDA_OK(cudaGetDevice(&c->device));
  void* p = nullptr;
  B200_CUDA_OK(cudaMalloc(&p, (size_t)ar_region_bytes(max_bytes)));
  B200_CUDA_OK(cudaMemset(p, 0, (size_t)ar_region_bytes(max_bytes)));
  B200_CUDA_OK(cudaDeviceSynchronize());
  c->local = static_cast<uint8_t*>(p);
  c->peer[rank] = c->local;
  memset(handle_out, 0, B200_AR_HANDLE_BYTES);
  if (world_size > 1) {
    cudaIpcMemHandle_t h;
    B200_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) <= B200_AR_HANDLE_BYTES, "handle blob too small");
    memcpy(handle_out, &h, sizeof(h));
  }
  *comm = c;
  return B200_OK;
}

int b200_ar_open_peers(b200_ar_comm* c, const void* all_handles) {
  B200_CHECK_ARG(c && all_handles, "ar_open_peers: null pointer");
  if (c->opened) return B200_OK;
  const uint8_t* hs = static_cast<const uint8_t*>(all_handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)r * B200_AR_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    B200_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = static_cast<uint8_t*>(p);
  }
  c->opened = true;
  return B200_OK;
}

static int ar_alloc_region(b200_ar_comm* c) {
  void* p = nullptr;
  B200_CUDA_OK(cudaMalloc(&p, (size_t)ar_region_bytes(c->max_bytes)));
  c->local = static_cast<uint8_t*>(p);
  c->peer[c->rank] = c->local;
  B200_CUDA_OK(cudaMemset(p, 0, (size_t)ar_region_bytes(c->max_bytes)));
  B200_CUDA_OK(cudaDeviceSynchronize());
  return B200_OK;
}

static int ar_create_all(b200_ar_comm** comms, const int* devices, int world, int64_t max_bytes) {
  for (int r = 0; r < world; ++r) {
    B200_CUDA_OK(cudaSetDevice(devices[r]));
    auto* c = comms[r] = new b200_ar_comm();
    c->rank = r;
    c->world = world;
    c->max_bytes = max_bytes;
    c->device = devices[r];
    c->ipc = false;
    const int rc = ar_alloc_region(c);
    if (rc != B200_OK) return rc;
  }
  for (int r = 0; r < world; ++r) {
    B200_CUDA_OK(cudaSetDevice(devices[r]));
    for (int q = 0; q < world; ++q) {
      if (q == r) continue;
      int can = 0;
      B200_CUDA_OK(cudaDeviceCanAccessPeer(&can, devices[r], devices[q]));
      if (!can)
        return set_error(B200_ERR_UNSUPPORTED, "ar_create_all: device %d cannot access device %d",
                         devices[r], devices[q]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled)
        (void)cudaGetLastError();  // enabled earlier (by torch or a previous group): fine
      else
        B200_CUDA_OK(e);
      comms[r]->peer[q] = comms[q]->local;
    }
    comms[r]->opened = true;
  }
  return B200_OK;
}

int b200_ar_create_all(b200_ar_comm** comms, const int* devices, int world_size, int64_t max_bytes) {
  B200_CHECK_ARG(comms && devices, "ar_create_all: null pointer");
  B200_CHECK_ARG(world_size >= 1 && world_size <= AR_MAX_WORLD, "ar_create_all: world %d unsupported (max %d)",
                 world_size, AR_MAX_WORLD);
  B200_CHECK_ARG(max_bytes > 0 && max_bytes % 16 == 0, "ar_create_all: max_bytes must be a multiple of 16");
  for (int r = 0; r < world_size; ++r) {
    comms[r] = nullptr;
    for (int q = 0; q < r; ++q)
      B200_CHECK_ARG(devices[q] != devices[r], "ar_create_all: device %d listed twice", devices[r]);
  }
  int prev = 0;
  B200_CUDA_OK(cudaGetDevice(&prev));
  const int rc = ar_create_all(comms, devices, world_size, max_bytes);
  if (rc != B200_OK) {  // all or nothing
    for (int r = 0; r < world_size; ++r) {
      if (comms[r]) {
        cudaSetDevice(comms[r]->device);
        if (comms[r]->local) cudaFree(comms[r]->local);
        delete comms[r];
        comms[r] = nullptr;
      }
    }
  }
  cudaSetDevice(prev);
  return rc;
}

static int ar_launch(b200_ar_comm* c, void* data, int64_t count, int dtype, const float* partials,
                     const W4Plan& plan, int64_t row_n, b200_stream_t stream);

int b200_ar_allreduce(b200_ar_comm* c, void* data, int64_t count, int dtype,
                      b200_stream_t stream) {
  return ar_launch(c, data, count, dtype, nullptr, W4Plan{}, 0, stream);
}

int b200_ar_allreduce_splitk(b200_ar_comm* c, void* out, const float* partials, int splits,
                             int64_t gemm_k, int64_t n, int64_t count, int dtype,
                             b200_stream_t stream) {
  B200_CHECK_ARG(partials && n > 0 && n % 128 == 0 && gemm_k > 0 && gemm_k % 128 == 0 &&
                     count % n == 0,
                 "ar_allreduce_splitk: partials of a [K, N] GEMM: n %% 128 == 0, count %% n == 0");
  const W4Plan plan = w4_get_plan(n, gemm_k, count / n);
  B200_CHECK_ARG(splits == plan.slots, "ar_allreduce_splitk: expected %d partial slots, got %d",
                 plan.slots, splits);
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "ar_allreduce_splitk: bf16 / fp16 output");
  B200_CHECK_ARG(c && c->world > 1, "ar_allreduce_splitk: needs world_size > 1");
  return ar_launch(c, out, count, dtype, partials, plan, n, stream);
}

static int ar_launch(b200_ar_comm* c, void* data, int64_t count, int dtype, const float* partials,
                     const W4Plan& plan, int64_t row_n, b200_stream_t stream) {
  B200_CHECK_ARG(c && data, "ar_allreduce: null pointer");
  B200_CHECK_ARG(dtype >= 0 && dtype <= 2, "ar_allreduce: bad dtype");
  if (count == 0 || c->world == 1) return B200_OK;
  B200_CHECK_ARG(c->opened, "ar_allreduce: peers not opened");
  const int es = dtype == B200_FP32 ? 4 : 2;
  const int64_t bytes = count * es;
  B200_CHECK_ARG(bytes % 16 == 0 && is_aligned(data, 16),
                 "ar_allreduce: message must be 16-byte aligned and a multiple of 16 bytes");
  if (bytes > c->max_bytes)
    return set_error(B200_ERR_WORKSPACE, "ar_allreduce: %lld B exceeds the %lld B symmetric buffer",
                     (long long)bytes, (long long)c->max_bytes);
  const int64_t nvec = bytes / 16;
  int blocks = (int)((nvec + 2 * AR_THREADS - 1) / (2 * AR_THREADS));
  if (blocks < 1) blocks = 1;
  if (blocks > AR_MAX_BLOCKS) blocks = AR_MAX_BLOCKS;
  ArDevPtrs ptrs{};
  for (int r = 0; r < c->world; ++r) ptrs.base[r] = c->peer[r];
  auto st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case B200_BF16:
      allreduce_oneshot_kernel<__nv_bfloat16, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<__nv_bfloat16*>(data), nvec, c->rank, c->world, c->max_bytes, partials,
          plan, (int)row_n, count, ArNormArgs<__nv_bfloat16>{});
      break;
    case B200_FP16:
      allreduce_oneshot_kernel<__half, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<__half*>(data), nvec, c->rank, c->world, c->max_bytes, partials,
          plan, (int)row_n, count, ArNormArgs<__half>{});
      break;
    default:
      allreduce_oneshot_kernel<float, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<float*>(data), nvec, c->rank, c->world, c->max_bytes, nullptr, W4Plan{}, 0, 0,
          ArNormArgs<float>{});
      break;
  }
  B200_LAUNCH_OK("allreduce_oneshot");
  return B200_OK;
}

int b200_ar_allreduce_splitk_norm(b200_ar_comm* c, void* out, void* residual, const float* partials,
                                  int splits, int64_t gemm_k, const void* weight, int64_t rows,
                                  int64_t n, float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(c && out && residual && partials && weight, "ar_allreduce_splitk_norm: null pointer");
  B200_CHECK_ARG(c->world > 1 && c->opened, "ar_allreduce_splitk_norm: needs an opened communicator, world > 1");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "ar_allreduce_splitk_norm: bf16 / fp16 only");
  B200_CHECK_ARG(rows >= 1 && rows <= AR_MAX_BLOCKS && n > 0 && n % 128 == 0 && n / 8 <= AR_THREADS &&
                     gemm_k > 0 && gemm_k % 128 == 0,
                 "ar_allreduce_splitk_norm: rows <= %d, n %% 128 == 0, n <= %d", AR_MAX_BLOCKS,
                 AR_THREADS * 8);
  B200_CHECK_ARG(is_aligned(out, 16) && is_aligned(residual, 16) && is_aligned(weight, 16) &&
                     is_aligned(partials, 16),
                 "ar_allreduce_splitk_norm: 16-byte alignment required");
  const int64_t count = rows * n, bytes = count * 2;
  if (bytes > c->max_bytes)
    return set_error(B200_ERR_WORKSPACE, "ar_allreduce_splitk_norm: %lld B exceeds the %lld B symmetric buffer",
                     (long long)bytes, (long long)c->max_bytes);
  const W4Plan plan = w4_get_plan(n, gemm_k, rows);
  B200_CHECK_ARG(splits == plan.slots, "ar_allreduce_splitk_norm: expected %d partial slots, got %d",
                 plan.slots, splits);
  ArDevPtrs ptrs{};
  for (int r = 0; r < c->world; ++r) ptrs.base[r] = c->peer[r];
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t nvec = count / 8;
  const int blocks = (int)rows;  // one block per row: slice = ceil(nvec / blocks) = n / 8 vectors
  if (dtype == B200_BF16) {
    ArNormArgs<__nv_bfloat16> na{static_cast<__nv_bfloat16*>(residual),
                                 static_cast<const __nv_bfloat16*>(weight),
                                 static_cast<__nv_bfloat16*>(out), eps};
    allreduce_oneshot_kernel<__nv_bfloat16, true><<<blocks, AR_THREADS, 0, st>>>(
        ptrs, static_cast<__nv_bfloat16*>(nullptr), nvec, c->rank, c->world, c->max_bytes, partials,
        plan, (int)n, count, na);
  } else {
    ArNormArgs<__half> na{static_cast<__half*>(residual), static_cast<const __half*>(weight),
                          static_cast<__half*>(out), eps};
    allreduce_oneshot_kernel<__half, true><<<blocks, AR_THREADS, 0, st>>>(
        ptrs, static_cast<__half*>(nullptr), nvec, c->rank, c->world, c->max_bytes, partials, plan,
        (int)n, count, na);
  }
  B200_LAUNCH_OK("allreduce_oneshot_norm");
  return B200_OK;
}

int b200_ar_destroy(b200_ar_comm* c) {
  if (!c) return B200_OK;
  if (c->ipc) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
  } else if (c->local) {  // same-process group: the region lives on the communicator's own device
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaFree(c->local);
    cudaSetDevice(prev);
  }
  delete c;
  return B200_OK;
}

}  // extern "C"
