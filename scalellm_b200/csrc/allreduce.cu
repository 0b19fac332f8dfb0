// allreduce.cu — tensor-parallel collectives over NVLink peer memory.
//
// Replaces ProcessGroupNCCL::allreduce / allgather (src/model_parallel/process_group.cpp:135-202)
// for the latency-bound <= 1 MiB collectives of the decode step (2 all-reduces per layer: after
// o_proj and down_proj; the embedding / lm_head gathers).
//
// One process per GPU (b200_ar_create + b200_ar_open_peers, CUDA IPC), or all ranks in one
// process with one thread per GPU like the reference's engine (b200_ar_create_all, the
// counterpart of its ncclCommInitAll, process_group.cpp:98-118: peer access instead of IPC).
// Each rank owns ONE cudaMalloc'ed symmetric region
//   [ LL inbox p0 | p1 | LL result p0 | p1 | staging p0 | p1 | flags1[world][128] | flags2[128] |
//     flagsA[world][128] | cand[2][world][128] | epoch | done ]
// exported through CUDA IPC and mapped by every peer (NVSwitch gives every pair full NVLink
// bandwidth).  Every collective of a communicator takes the next epoch e (kept in device memory
// and advanced by the kernel, so a captured CUDA graph replays) and uses the buffers of parity
// e & 1; flags carry epoch values, so nothing is ever reset.  Reuse is safe because a rank can be
// in call e+2 only after every rank has left call e: finishing call e+1 needs a flag that some
// owner publishes only after it has seen every rank's call-e+1 contribution.
//
// Two algorithms for the all-reduce (B200_AR_ALGO = oneshot | twoshot; default twoshot above two ranks):
//   * ONE-SHOT (world 2): every block stages its slice in its own buffer, publishes a flag to every
//     peer, waits for theirs and pulls all `world` copies (all loads in flight before the first add).
//   * TWO-SHOT, column partitioned (the form that scales to 8 ranks): the message is cut into rows
//     (one block per row; for the decoder's [tokens, hidden] messages a row is a token) and every
//     row into `world` column chunks; chunk o of every row is owned by rank o.  Every rank PUSHES
//     its contribution of a chunk into the owner's inbox (slot = source rank); the owner sums the
//     world slots in rank order in fp32 (bit-identical to the one-shot result and to a host sum in
//     rank order), rounds once and pushes the reduced chunk into every rank's result buffer; every
//     rank then consumes the row from its own memory.  Every block of every rank sends and reduces
//     the same amount (the first cut of this round partitioned by rows: 8 of a rank's 64 blocks did
//     all of its reducing and 7x the sending, 15.9 us per fused call at 8 ranks against 9.6 at 2).
//     Per rank and call NVLink carries (world-1)/world of the message out and the same in, twice —
//     0.9 MB at 8 ranks instead of the one-shot's 3.7 MB pull — and the critical path is two
//     one-way NVLink hops, not `world` dependent round trips.
//
// Fused forms (the row-parallel GEMM -> all-reduce -> residual add -> RMSNorm chain of a TP
// decoder layer, models/meta/llama.h:170-177):
//   * b200_ar_allreduce_splitk: the contribution is read from the producing W4A16 GEMM's fp32
//     stream-K partials, summing each tile's contributor slots (common.cuh W4Plan), instead of a
//     bf16 input;
//   * b200_ar_allreduce_splitk_norm: additionally the reduced row is consumed in place:
//     residual += x; out = rms_norm(residual) * w — one launch for the whole chain, bit-identical
//     to the separate kernels (same rounding points).
// b200_ar_argmax: greedy sampling over a vocabulary-sharded lm_head without gathering logits:
// local argmax per row, an 8-byte candidate per rank and row exchanged through `cand`, global
// winner picked with torch.argmax's tie rule (lm_head gather + argmax: llama.h:259-265, sampler).

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace b200 {
constexpr int AR_MAX_WORLD = 8;
constexpr int AR_MAX_BLOCKS = 64;   // one-shot kernels: blocks (flag slots per source rank)
constexpr int AR_MAX_ROWS = 128;    // two-shot kernels: rows (= blocks) per call
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

// region layout (byte offsets)
__host__ __device__ inline int64_t ar_result_off(int64_t max_bytes) { return 2 * max_bytes; }
// staging buffers (2 parities) of the flag-synchronised kernels (one-shot all-reduce, all-gather):
// kept apart from the LL buffers, whose stale contents must never look like a current epoch
__host__ __device__ inline int64_t ar_stage_off(int64_t max_bytes) { return 4 * max_bytes; }
__host__ __device__ inline int64_t ar_flags_off(int64_t max_bytes) { return 6 * max_bytes; }  // flags1
__host__ __device__ inline int64_t ar_flags2_off(int64_t max_bytes) {
  return ar_flags_off(max_bytes) + (int64_t)AR_MAX_WORLD * AR_MAX_ROWS * 4;
}
__host__ __device__ inline int64_t ar_flagsA_off(int64_t max_bytes) {
  return ar_flags2_off(max_bytes) + (int64_t)AR_MAX_ROWS * 4;
}
__host__ __device__ inline int64_t ar_cand_off(int64_t max_bytes) {
  return ar_flagsA_off(max_bytes) + (int64_t)AR_MAX_WORLD * AR_MAX_ROWS * 4;
}
__host__ __device__ inline int64_t ar_epoch_off(int64_t max_bytes) {
  return ar_cand_off(max_bytes) + 2ll * AR_MAX_WORLD * AR_MAX_ROWS * 8;
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
__device__ __forceinline__ void ar_wait_flag(const uint32_t* p, uint32_t e) {
  while ((int32_t)(ld_acquire_sys(p) - e) < 0) {
  }
}
// debug: %globaltimer (ns, common to all SMs of a GPU; GPUs of one box agree to ~1 us) into
// trace[block * 8 + slot] when b200_debug_set_trace installed a buffer (tools/ar_bench.py)
__device__ __forceinline__ void ar_stamp(long long* trace, int slot) {
  if (trace != nullptr && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    trace[(int64_t)blockIdx.x * 8 + slot] = (long long)ns;
  }
}
// last block out advances the epoch (device-side state: graph replay safe)
__device__ __forceinline__ void ar_finish(uint32_t* epoch_ptr, uint32_t e) {
  uint32_t* done_ptr = epoch_ptr + 1;
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

template <typename T>
__device__ __forceinline__ void accum16(float (&acc)[16 / sizeof(T)], uint4 v) {
  const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(16 / sizeof(T)); ++i) acc[i] += Num<T>::to_f(e[i]);
}

// acc = sum over ranks 0..world-1 (that order, fp32) of vector i of every rank's buffer at
// `off`: all loads are issued before the first add, so the NVLink round trips overlap.
template <typename T>
__device__ __forceinline__ void ar_pull_sum(float (&acc)[16 / sizeof(T)], const ArDevPtrs& ptrs,
                                            int64_t off, int64_t i, int world) {
  uint4 v[AR_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < AR_MAX_WORLD; ++r)
    if (r < world) v[r] = ld_volatile_v4(reinterpret_cast<const uint4*>(ptrs.base[r] + off) + i);
#pragma unroll
  for (int k = 0; k < (int)(16 / sizeof(T)); ++k) acc[k] = 0.f;
#pragma unroll
  for (int r = 0; r < AR_MAX_WORLD; ++r)
    if (r < world) accum16<T>(acc, v[r]);
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
                                                                      ArNormArgs<T> na,
                                                                      long long* trace) {
  constexpr int VEC = 16 / sizeof(T);
  ar_stamp(trace, 0);
  uint8_t* local = ptrs.base[rank];
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(local + ar_epoch_off(max_bytes));
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int64_t buf_off = ar_stage_off(max_bytes) + ((e & 1) ? max_bytes : 0);

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
  ar_stamp(trace, 1);

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
  ar_stamp(trace, 2);

  // 4. reduce in rank order, write back in place (or feed the residual + RMSNorm epilogue)
  if constexpr (!NORM) {
    for (int64_t i = v0 + threadIdx.x; i < v1; i += AR_THREADS) {
      float acc[VEC];
      ar_pull_sum<T>(acc, ptrs, buf_off, i, world);
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int k = 0; k < VEC; ++k) oe[k] = Num<T>::from_f(acc[k]);
      reinterpret_cast<uint4*>(data)[i] = o;
    }
  } else {
    __shared__ float red[32];
    extern __shared__ float sq[];        // [row_n] fp32: the row, for the reference-ordered sum of squares
    const int64_t i = v0 + threadIdx.x;  // one vector per thread: the block's slice is one row
    const bool have = i < v1;
    float x[VEC];
    if (have) {
      float acc[VEC];
      ar_pull_sum<T>(acc, ptrs, buf_off, i, world);
      uint4 rraw = reinterpret_cast<const uint4*>(na.residual)[i];
      const T* rr = reinterpret_cast<const T*>(&rraw);
      uint4 sraw;
      T* sv = reinterpret_cast<T*>(&sraw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float reduced = rnd<T>(acc[k]);  // what the plain all-reduce would have stored
        const float f = Num<T>::to_f(rr[k]) + reduced;
        sq[threadIdx.x * VEC + k] = f;
        sv[k] = Num<T>::from_f(f);
        x[k] = Num<T>::to_f(sv[k]);
      }
      reinterpret_cast<uint4*>(na.residual)[i] = sraw;
    }
    const float total = row_sumsq_ref_order<AR_THREADS>(sq, red, row_n);
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

  ar_stamp(trace, 5);
  ar_finish(epoch_ptr, e);
}

// ---------------------------------------------------------------------------------------------
// Two-shot, column-partitioned all-reduce (file header), LL ("low latency") transport: every 16-byte
// store that crosses NVLink carries 8 bytes of payload and two copies of the call's epoch —
// { data0, epoch, data1, epoch } — and the receiver polls the data itself until both epoch words
// match (8-byte aligned stores are single-copy atomic, so a matching flag vouches for its
// payload).  No flag arrays, no release fences, no __syncthreads on the exchange path: the two hops
// cost two one-way NVLink latencies instead of two (remote write + system fence + flag + poll)
// sequences — the flag form measured 15-17 us per fused call at two ranks with both ranks in
// lock step (profiles/r02_ar_bench.md).  The price is 2x the bytes on the wire and in the buffers
// (rows * row_bytes * 2 <= max_bytes), irrelevant at <= 1 MiB.
//
// One block per row; a row is `row_vecs` 16-byte payload vectors (the last row of a plain message
// may be shorter), thread t holds vectors t, t + 512, ... (VPT of them).  Contribution source:
//   FROM_PARTIALS: the producing GEMM's stream-K partials (sum of the tile's slots, rounded once);
//   else:          data_in (T).
// NORM: the reduced row feeds residual += x; out = rms_norm(residual) * weight (ArNormArgs);
// else it is stored to data_out.  T = float only without partials / norm.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint4* dst, uint32_t d0, uint32_t d1, uint32_t e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(d0), "r"(e), "r"(d1), "r"(e)
               : "memory");
}
// payload vector (16 B) -> two LL lines at ll[2 * v], ll[2 * v + 1]
__device__ __forceinline__ void ll_store_vec(uint4* ll, int64_t v, uint4 p, uint32_t e) {
  ll_store(ll + 2 * v, p.x, p.y, e);
  ll_store(ll + 2 * v + 1, p.z, p.w, e);
}
__device__ __forceinline__ uint4 ll_load_vec(const uint4* ll, int64_t v, uint32_t e) {
  uint4 a, b;
  do {
    a = ld_volatile_v4(ll + 2 * v);
    b = ld_volatile_v4(ll + 2 * v + 1);
  } while (a.y != e || a.w != e || b.y != e || b.w != e);
  return make_uint4(a.x, a.z, b.x, b.z);
}

template <typename T, int VPT, bool FROM_PARTIALS, bool NORM>
__global__ void __launch_bounds__(AR_THREADS) allreduce_twoshot_kernel(
    ArDevPtrs ptrs, const T* __restrict__ data_in, T* __restrict__ data_out, int64_t nvec_total,
    int row_vecs, int rank, int world, int64_t max_bytes, const float* __restrict__ partials,
    W4Plan plan, int row_n, int64_t split_stride, ArNormArgs<T> na, long long* trace) {
  constexpr int VEC = 16 / sizeof(T);
  ar_stamp(trace, 6);
  pdl_launch_dependents();  // the next GEMM may start prefetching its weights while we exchange
  // Before griddepcontrol.wait: whatever does not depend on the producing GEMM.  In the fused (NORM)
  // form the producer is always a GEMM, which triggers its dependents after its own wait: when this
  // kernel starts, everything older than that GEMM — the previous norm (the residual stream) — is
  // complete and visible.  The epoch is NOT read here: a plain all-reduce may directly follow another
  // one of this kernel (sliced large messages), which triggers at its top, and would see the epoch
  // of the call still in flight (ranks then disagree on the epoch: a hang, found by tools/ar_bench.py).
  uint8_t* local = ptrs.base[rank];
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(local + ar_epoch_off(max_bytes));

  const int row = blockIdx.x, rows = gridDim.x;
  const int C = (row_vecs + world - 1) / world;  // vectors of a row per owner (column partition)
  const int64_t row_v0 = (int64_t)row * row_vecs;
  const int64_t left = nvec_total - row_v0;
  const int row_len = left < row_vecs ? (int)left : row_vecs;
  uint4 res_pre[VPT], w_pre[VPT];
  int cnt_pre[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int j = threadIdx.x + k * AR_THREADS;
    if (j < row_len) {
      if constexpr (NORM) {
        res_pre[k] = reinterpret_cast<const uint4*>(na.residual)[row_v0 + j];
        w_pre[k] = *reinterpret_cast<const uint4*>(na.weight + j * VEC);
      }
      if constexpr (FROM_PARTIALS) cnt_pre[k] = w4_contrib_col(plan, j * 8);
    }
  }
  pdl_wait();               // the contribution comes from the producing GEMM
  // the epoch of this call: its load is in flight together with the partial-sum loads below
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int64_t par_off = (e & 1) ? max_bytes : 0;
  ar_stamp(trace, 0);

  // ---- 1. push my contribution of every vector into its owner's inbox, slot = my rank ----
  {
    const int64_t slot_v0 = ((int64_t)rank * rows + row) * C;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = threadIdx.x + k * AR_THREADS;
      if (j < row_len) {
        const int owner = j / C;
        uint4* inbox = reinterpret_cast<uint4*>(ptrs.base[owner] + par_off);
        uint4 c;
        if constexpr (FROM_PARTIALS) {
          float a[8];
          // 8 consecutive columns of this row, inside one n tile
          w4_sum_partials8(a, partials + (row_v0 + j) * 8, split_stride, cnt_pre[k]);
          T* ce = reinterpret_cast<T*>(&c);
#pragma unroll
          for (int q = 0; q < 8; ++q) ce[q] = Num<T>::from_f(a[q]);
        } else {
          c = reinterpret_cast<const uint4*>(data_in)[row_v0 + j];
        }
        ll_store_vec(inbox, slot_v0 + (j - owner * C), c, e);
      }
    }
  }
  ar_stamp(trace, 1);

  // ---- 2. the owner of a vector reduces its world slots in rank order and pushes the result to
  //         everyone; 3. everyone else picks the reduced vector up from its own result buffer ----
  uint4 red[VPT];
  {
    const uint4* inbox = reinterpret_cast<const uint4*>(local + par_off);
    const uint4* res = reinterpret_cast<const uint4*>(local + ar_result_off(max_bytes) + par_off);
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = threadIdx.x + k * AR_THREADS;
      if (j >= row_len) continue;
      const int owner = j / C;
      if (owner == rank) {
        uint4 v[AR_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r)
          if (r < world) v[r] = ll_load_vec(inbox, ((int64_t)r * rows + row) * C + (j - owner * C), e);
        float acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r)
          if (r < world) accum16<T>(acc, v[r]);
        T* oe = reinterpret_cast<T*>(&red[k]);
#pragma unroll
        for (int q = 0; q < VEC; ++q) oe[q] = Num<T>::from_f(acc[q]);
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r)
          if (r < world && r != rank)
            ll_store_vec(reinterpret_cast<uint4*>(ptrs.base[r] + ar_result_off(max_bytes) + par_off),
                         row_v0 + j, red[k], e);
      } else {
        red[k] = ll_load_vec(res, row_v0 + j, e);
      }
    }
    ar_stamp(trace, 4);
  }

  if constexpr (!NORM) {
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = threadIdx.x + k * AR_THREADS;
      if (j < row_len) reinterpret_cast<uint4*>(data_out)[row_v0 + j] = red[k];
    }
  } else {
    __shared__ float redsm[32];
    extern __shared__ float sq[];  // [row_n] fp32: the row, for the reference-ordered sum of squares
    float x[VPT][VEC];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = threadIdx.x + k * AR_THREADS;
      if (j < row_len) {
        const T* rv = reinterpret_cast<const T*>(&red[k]);  // what the plain all-reduce would store
        const T* rr = reinterpret_cast<const T*>(&res_pre[k]);
        uint4 sraw;
        T* sv = reinterpret_cast<T*>(&sraw);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const float f = Num<T>::to_f(rr[q]) + Num<T>::to_f(rv[q]);
          sq[j * VEC + q] = f;
          sv[q] = Num<T>::from_f(f);
          x[k][q] = Num<T>::to_f(sv[q]);
        }
        reinterpret_cast<uint4*>(na.residual)[row_v0 + j] = sraw;
      }
    }
    const float total = row_sumsq_ref_order<AR_THREADS>(sq, redsm, row_n);
    const float rstd = rsqrtf(total / row_n + na.eps);
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = threadIdx.x + k * AR_THREADS;
      if (j < row_len) {
        const T* w = reinterpret_cast<const T*>(&w_pre[k]);
        uint4 oraw;
        T* o = reinterpret_cast<T*>(&oraw);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const float y = rnd<T>(x[k][q] * rstd);
          o[q] = Num<T>::from_f(y * Num<T>::to_f(w[q]));
        }
        reinterpret_cast<uint4*>(na.out)[row_v0 + j] = oraw;
      }
    }
  }
  ar_stamp(trace, 5);
  ar_finish(epoch_ptr, e);
}

// ---------------------------------------------------------------------------------------------
// Greedy sampling over a vocabulary-sharded lm_head: out[row] = argmax over ALL ranks' columns of
// logits[row, :], as torch.argmax(cat(all-gather(logits), -1)) would give (first index of the
// maximum, NaN counts as the maximum) — without moving the logits.  One block per row: local
// argmax of this rank's [n_local] shard, (value, global index) pushed to every rank's `cand`
// table, flagsA[src][row] published, then the world candidates are compared locally.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ar_argmax_better(float a, int64_t ia, float b, int64_t ib) {
  const bool an = a != a, bn = b != b;
  if (an != bn) return an;
  if (an) return ia < ib;
  return a > b || (a == b && ia < ib);
}

template <typename T>
__global__ void __launch_bounds__(AR_THREADS) argmax_sharded_kernel(
    ArDevPtrs ptrs, int64_t* __restrict__ out, const T* __restrict__ x, int n, int64_t stride,
    int rank, int world, int64_t max_bytes) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float sv[16];
  __shared__ int si[16];
  pdl_wait();
  pdl_launch_dependents();
  uint8_t* local = ptrs.base[rank];
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(local + ar_epoch_off(max_bytes));
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int par = e & 1;
  const int row = blockIdx.x;
  const T* xr = x + (int64_t)row * stride;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const bool vec = (stride % VEC == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int nv = vec ? n / VEC : 0;
  for (int v = threadIdx.x; v < nv; v += AR_THREADS) {
    const uint4 raw = ld_nc_v4(xr + v * VEC);
    const T* el = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float f = Num<T>::to_f(el[i]);
      if (ar_argmax_better(f, v * VEC + i, best, bi)) { best = f; bi = v * VEC + i; }
    }
  }
  for (int j = nv * VEC + threadIdx.x; j < n; j += AR_THREADS) {
    const float f = Num<T>::to_f(xr[j]);
    if (ar_argmax_better(f, j, best, bi)) { best = f; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ar_argmax_better(ob, oi, best, bi)) { best = ob; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < AR_THREADS / 32 ? sv[lane] : -INFINITY;
    bi = lane < AR_THREADS / 32 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ar_argmax_better(ob, oi, best, bi)) { best = ob; bi = oi; }
    }
    // every lane holds the local winner; lane r tells rank r, then collects rank r's candidate
    const int gi = bi + rank * n;  // global column (shards are equal, rank-major: cat order)
    const uint64_t mine = ((uint64_t)(uint32_t)gi << 32) | __float_as_uint(best);
    float cv = -INFINITY;
    int64_t ci = INT64_MAX;
    if (lane < world) {
      uint64_t* cand = reinterpret_cast<uint64_t*>(ptrs.base[lane] + ar_cand_off(max_bytes));
      cand[((int64_t)par * AR_MAX_WORLD + rank) * AR_MAX_ROWS + row] = mine;
      uint32_t* fa = reinterpret_cast<uint32_t*>(ptrs.base[lane] + ar_flagsA_off(max_bytes));
      st_release_sys(&fa[rank * AR_MAX_ROWS + row], e);
      const uint32_t* my_fa = reinterpret_cast<const uint32_t*>(local + ar_flagsA_off(max_bytes));
      ar_wait_flag(&my_fa[lane * AR_MAX_ROWS + row], e);
      const volatile uint64_t* my_cand =
          reinterpret_cast<const volatile uint64_t*>(local + ar_cand_off(max_bytes));
      const uint64_t c = my_cand[((int64_t)par * AR_MAX_WORLD + lane) * AR_MAX_ROWS + row];
      cv = __uint_as_float((uint32_t)c);
      ci = (int64_t)(uint32_t)(c >> 32);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (ar_argmax_better(ov, oi, cv, ci)) { cv = ov; ci = oi; }
    }
    if (lane == 0) out[row] = ci;
  }
  ar_finish(epoch_ptr, e);
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
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_ptr) + 1;
  const int64_t buf_off = ar_stage_off(max_bytes) + ((e & 1) ? max_bytes : 0);
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
    uint4 v[AR_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < AR_MAX_WORLD; ++r)
      if (r < world) v[r] = ld_volatile_v4(reinterpret_cast<const uint4*>(ptrs.base[r] + buf_off) + i);
#pragma unroll
    for (int r = 0; r < AR_MAX_WORLD; ++r)
      if (r < world) out[(row * world + r) * cols_v + c] = v[r];
  }
  ar_finish(epoch_ptr, e);
}

}  // namespace b200

using namespace b200;

// B200_AR_ALGO = oneshot | twoshot (read once); default: two-shot (LL transport)
static bool ar_use_twoshot(int world) {
  static const int forced = [] {
    const char* v = getenv("B200_AR_ALGO");
    if (v && !strcmp(v, "oneshot")) return 1;
    if (v && !strcmp(v, "twoshot")) return 2;
    return 0;
  }();
  if (forced) return forced == 2;
  return world >= 2;
}

static ArDevPtrs ar_ptrs(const b200_ar_comm* c) {
  ArDevPtrs ptrs{};
  for (int r = 0; r < c->world; ++r) ptrs.base[r] = c->peer[r];
  return ptrs;
}

// two-shot launch: `rows` rows of `row_vecs` 16-byte vectors (the last may be short: nvec total)
template <typename T, bool FROM_PARTIALS, bool NORM>
static int ar_launch_twoshot(b200_ar_comm* c, const T* in, T* out, int64_t nvec, int row_vecs,
                             int rows, const float* partials, const W4Plan& plan, int row_n,
                             int64_t split_stride, ArNormArgs<T> na, cudaStream_t st) {
  const int C = (row_vecs + c->world - 1) / c->world;        // vectors of a row per owner
  if ((int64_t)c->world * C * rows * 32 > c->max_bytes)      // LL lines: 2x the payload bytes
    return set_error(B200_ERR_WORKSPACE, "ar_allreduce: %d rows of %d B exceed the %lld B symmetric buffer",
                     rows, row_vecs * 16, (long long)c->max_bytes);
  const ArDevPtrs ptrs = ar_ptrs(c);
  const size_t sm = NORM ? (size_t)row_n * 4 : 0;  // the row in fp32 (<= 32 KB)
  if (row_vecs <= AR_THREADS) {
    B200_PDL_LAUNCH_L(1, "allreduce_twoshot", (allreduce_twoshot_kernel<T, 1, FROM_PARTIALS, NORM>),
                      (unsigned)rows, AR_THREADS, sm, st, ptrs, in, out, nvec, row_vecs, c->rank, c->world,
                      c->max_bytes, partials, plan, row_n, split_stride, na, debug_trace_ptr());
  } else {
    B200_PDL_LAUNCH_L(1, "allreduce_twoshot", (allreduce_twoshot_kernel<T, 2, FROM_PARTIALS, NORM>),
                      (unsigned)rows, AR_THREADS, sm, st, ptrs, in, out, nvec, row_vecs, c->rank, c->world,
                      c->max_bytes, partials, plan, row_n, split_stride, na, debug_trace_ptr());
  }
  return B200_OK;
}

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
  const ArDevPtrs ptrs = ar_ptrs(c);
  allgather_oneshot_kernel<<<blocks, AR_THREADS, 0, st>>>(
      ptrs, static_cast<const uint4*>(in), static_cast<uint4*>(out), nvec, row_bytes / 16, c->rank,
      c->world, c->max_bytes);
  B200_LAUNCH_OK("allgather_oneshot");
  return B200_OK;
}

static int ar_create_one(b200_ar_comm* c, void* handle_out) {
  B200_CUDA_OK(cudaGetDevice(&c->device));
  void* p = nullptr;
  B200_CUDA_OK(cudaMalloc(&p, (size_t)ar_region_bytes(c->max_bytes)));
  c->local = static_cast<uint8_t*>(p);
  c->peer[c->rank] = c->local;
  B200_CUDA_OK(cudaMemset(p, 0, (size_t)ar_region_bytes(c->max_bytes)));
  B200_CUDA_OK(cudaDeviceSynchronize());
  memset(handle_out, 0, B200_AR_HANDLE_BYTES);
  if (c->world > 1) {
    cudaIpcMemHandle_t h;
    B200_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) <= B200_AR_HANDLE_BYTES, "handle blob too small");
    memcpy(handle_out, &h, sizeof(h));
  }
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
  const int rc = ar_create_one(c, handle_out);
  if (rc != B200_OK) {  // nothing leaks on an early error
    if (c->local) cudaFree(c->local);
    delete c;
    *comm = nullptr;
    return rc;
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
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {  // close what was opened so far; the communicator stays unopened
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer[q]) {
          cudaIpcCloseMemHandle(c->peer[q]);
          c->peer[q] = nullptr;
        }
      return set_error(B200_ERR_CUDA, "ar_open_peers: cudaIpcOpenMemHandle(rank %d) failed: %s", r,
                       cudaGetErrorString(e));
    }
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
  auto st = static_cast<cudaStream_t>(stream);
  // two-shot: rows of the producing GEMM when there are partials, else 8 KiB (16 KiB) chunks
  int row_vecs = partials ? (int)(row_n * es / 16) : AR_THREADS;
  if (!partials && nvec > (int64_t)AR_THREADS * AR_MAX_ROWS) row_vecs = 2 * AR_THREADS;
  const int64_t rows = (nvec + row_vecs - 1) / row_vecs;
  if (ar_use_twoshot(c->world) && rows <= AR_MAX_ROWS && row_vecs <= 2 * AR_THREADS &&
      (int64_t)rows * (row_vecs + c->world) * 32 <= c->max_bytes) {
    switch (dtype) {
      case B200_BF16: {
        using T = __nv_bfloat16;
        return partials ? ar_launch_twoshot<T, true, false>(c, nullptr, static_cast<T*>(data), nvec, row_vecs, (int)rows, partials, plan, (int)row_n, count, ArNormArgs<T>{}, st)
                        : ar_launch_twoshot<T, false, false>(c, static_cast<const T*>(data), static_cast<T*>(data), nvec, row_vecs, (int)rows, nullptr, plan, 0, 0, ArNormArgs<T>{}, st);
      }
      case B200_FP16: {
        using T = __half;
        return partials ? ar_launch_twoshot<T, true, false>(c, nullptr, static_cast<T*>(data), nvec, row_vecs, (int)rows, partials, plan, (int)row_n, count, ArNormArgs<T>{}, st)
                        : ar_launch_twoshot<T, false, false>(c, static_cast<const T*>(data), static_cast<T*>(data), nvec, row_vecs, (int)rows, nullptr, plan, 0, 0, ArNormArgs<T>{}, st);
      }
      default:
        return ar_launch_twoshot<float, false, false>(c, static_cast<const float*>(data), static_cast<float*>(data), nvec, row_vecs, (int)rows, nullptr, plan, 0, 0, ArNormArgs<float>{}, st);
    }
  }
  int blocks = (int)((nvec + 2 * AR_THREADS - 1) / (2 * AR_THREADS));
  if (blocks < 1) blocks = 1;
  if (blocks > AR_MAX_BLOCKS) blocks = AR_MAX_BLOCKS;
  const ArDevPtrs ptrs = ar_ptrs(c);
  switch (dtype) {
    case B200_BF16:
      allreduce_oneshot_kernel<__nv_bfloat16, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<__nv_bfloat16*>(data), nvec, c->rank, c->world, c->max_bytes, partials,
          plan, (int)row_n, count, ArNormArgs<__nv_bfloat16>{}, debug_trace_ptr());
      break;
    case B200_FP16:
      allreduce_oneshot_kernel<__half, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<__half*>(data), nvec, c->rank, c->world, c->max_bytes, partials,
          plan, (int)row_n, count, ArNormArgs<__half>{}, debug_trace_ptr());
      break;
    default:
      allreduce_oneshot_kernel<float, false><<<blocks, AR_THREADS, 0, st>>>(
          ptrs, static_cast<float*>(data), nvec, c->rank, c->world, c->max_bytes, nullptr, W4Plan{}, 0, 0,
          ArNormArgs<float>{}, debug_trace_ptr());
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
  const bool twoshot = ar_use_twoshot(c->world);
  const int max_rows = twoshot ? AR_MAX_ROWS : AR_MAX_BLOCKS;
  const int max_n = (twoshot ? 2 : 1) * AR_THREADS * 8;
  B200_CHECK_ARG(rows >= 1 && rows <= max_rows && n > 0 && n % 128 == 0 && n <= max_n &&
                     gemm_k > 0 && gemm_k % 128 == 0,
                 "ar_allreduce_splitk_norm: rows <= %d, n %% 128 == 0, n <= %d", max_rows, max_n);
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
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t nvec = count / 8;
  if (twoshot) {
    if (dtype == B200_BF16) {
      using T = __nv_bfloat16;
      ArNormArgs<T> na{static_cast<T*>(residual), static_cast<const T*>(weight), static_cast<T*>(out), eps};
      return ar_launch_twoshot<T, true, true>(c, nullptr, nullptr, nvec, (int)(n / 8), (int)rows, partials,
                                              plan, (int)n, count, na, st);
    }
    using T = __half;
    ArNormArgs<T> na{static_cast<T*>(residual), static_cast<const T*>(weight), static_cast<T*>(out), eps};
    return ar_launch_twoshot<T, true, true>(c, nullptr, nullptr, nvec, (int)(n / 8), (int)rows, partials,
                                            plan, (int)n, count, na, st);
  }
  const ArDevPtrs ptrs = ar_ptrs(c);
  const int blocks = (int)rows;  // one block per row: slice = ceil(nvec / blocks) = n / 8 vectors
  if (dtype == B200_BF16) {
    ArNormArgs<__nv_bfloat16> na{static_cast<__nv_bfloat16*>(residual),
                                 static_cast<const __nv_bfloat16*>(weight),
                                 static_cast<__nv_bfloat16*>(out), eps};
    allreduce_oneshot_kernel<__nv_bfloat16, true><<<blocks, AR_THREADS, (size_t)n * 4, st>>>(
        ptrs, static_cast<__nv_bfloat16*>(nullptr), nvec, c->rank, c->world, c->max_bytes, partials,
        plan, (int)n, count, na, debug_trace_ptr());
  } else {
    ArNormArgs<__half> na{static_cast<__half*>(residual), static_cast<const __half*>(weight),
                          static_cast<__half*>(out), eps};
    allreduce_oneshot_kernel<__half, true><<<blocks, AR_THREADS, (size_t)n * 4, st>>>(
        ptrs, static_cast<__half*>(nullptr), nvec, c->rank, c->world, c->max_bytes, partials, plan,
        (int)n, count, na, debug_trace_ptr());
  }
  B200_LAUNCH_OK("allreduce_oneshot_norm");
  return B200_OK;
}

int b200_ar_argmax(b200_ar_comm* c, int64_t* out, const void* logits, int64_t rows, int64_t n_local,
                   int64_t stride, int dtype, b200_stream_t stream) {
  if (rows == 0) return B200_OK;  // an empty selection has no storage (every rank sees the same rows)
  B200_CHECK_ARG(c && out && logits, "ar_argmax: null pointer");
  B200_CHECK_ARG(rows >= 0 && rows <= AR_MAX_ROWS && n_local > 0 && stride >= n_local &&
                     n_local * (int64_t)c->world < (1ll << 31),
                 "ar_argmax: rows <= %d, world * n_local < 2^31", AR_MAX_ROWS);
  B200_CHECK_ARG(dtype >= 0 && dtype <= 2, "ar_argmax: bad dtype");
  if (rows == 0) return B200_OK;
  if (c->world == 1) return b200_argmax(out, logits, rows, n_local, stride, dtype, stream);
  B200_CHECK_ARG(c->opened, "ar_argmax: peers not opened");
  auto st = static_cast<cudaStream_t>(stream);
  const ArDevPtrs ptrs = ar_ptrs(c);
  switch (dtype) {
    case B200_BF16:
      B200_PDL_LAUNCH_L(1, "argmax_sharded", argmax_sharded_kernel<__nv_bfloat16>, (unsigned)rows, AR_THREADS, 0,
                        st, ptrs, out, static_cast<const __nv_bfloat16*>(logits), (int)n_local, stride,
                        c->rank, c->world, c->max_bytes);
      break;
    case B200_FP16:
      B200_PDL_LAUNCH_L(1, "argmax_sharded", argmax_sharded_kernel<__half>, (unsigned)rows, AR_THREADS, 0, st,
                        ptrs, out, static_cast<const __half*>(logits), (int)n_local, stride, c->rank,
                        c->world, c->max_bytes);
      break;
    default:
      B200_PDL_LAUNCH_L(1, "argmax_sharded", argmax_sharded_kernel<float>, (unsigned)rows, AR_THREADS, 0, st,
                        ptrs, out, static_cast<const float*>(logits), (int)n_local, stride, c->rank,
                        c->world, c->max_bytes);
      break;
  }
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
