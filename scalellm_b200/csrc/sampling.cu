// sampling.cu — the sampling tail's logits processors on the device:
// temperature, repetition / frequency / presence penalties and the in-place softmax, with the
// semantics (and rounding points) of the reference's kernels
//   src/kernels/sampling/penalty_kernels.cu:9-33,52-75,107-140   src/kernels/sampling/softmax_kernels.cu:11-54
// restated, never copied.  Greedy selection is b200_argmax / b200_ar_argmax (elementwise.cu,
// allreduce.cu).  All in place on logits [batch, vocab] (contiguous), one launch each, capturable.
#include <type_traits>

#include "common.cuh"

namespace b200 {

// logits[b, :] *= (t[b] == 0 ? 1 : 1 / t[b]) (penalty_kernels.cu:9-33).  The reference writes
// `logits[i] *= inv` with logits of type T and inv a float: c10's compound assignment converts the
// float to T FIRST (operator*=(T&, const T&)), so the inverse temperature is rounded to T, then the
// product is rounded to T.
template <typename T>
__global__ void __launch_bounds__(256) temperature_kernel(T* __restrict__ logits,
                                                          const T* __restrict__ temperatures,
                                                          int64_t batch, int64_t vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = batch * vocab;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float t = Num<T>::to_f(temperatures[i / vocab]);
    const float inv = rnd<T>(t == 0.f ? 1.0f : 1.0f / t);
    logits[i] = Num<T>::from_f(Num<T>::to_f(logits[i]) * inv);
  }
}

// token_ids [batch, max_len] int64 (unique ids per row, the first lens[b] are valid):
// logit < 0 ? logit * p : logit / p   (penalty_kernels.cu:52-75)
template <typename T>
__global__ void __launch_bounds__(256) repetition_penalty_kernel(
    T* __restrict__ logits, const int64_t* __restrict__ token_ids, const int32_t* __restrict__ lens,
    const T* __restrict__ penalties, int64_t max_len, int64_t vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t b = blockIdx.x;
  const float p = Num<T>::to_f(penalties[b]);
  T* row = logits + b * vocab;
  for (int i = threadIdx.x; i < lens[b]; i += blockDim.x) {
    const int64_t id = token_ids[b * max_len + i];
    if (id < 0 || id >= vocab) continue;  // the reference only asserts this in a comment
    const float x = Num<T>::to_f(row[id]);
    row[id] = Num<T>::from_f(x < 0.0f ? x * p : x / p);
  }
}

// logit -= count * freq; logit -= presence   for every token with count > 0 (penalty_kernels.cu:107-140)
template <typename T>
__global__ void __launch_bounds__(256) frequency_presence_penalty_kernel(
    T* __restrict__ logits, const int64_t* __restrict__ token_ids, const int32_t* __restrict__ counts,
    const int32_t* __restrict__ lens, const T* __restrict__ freq, const T* __restrict__ pres,
    int64_t max_len, int64_t vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t b = blockIdx.x;
  T* row = logits + b * vocab;
  const float f = Num<T>::to_f(freq[b]), pz = Num<T>::to_f(pres[b]);
  for (int i = threadIdx.x; i < lens[b]; i += blockDim.x) {
    const int64_t id = token_ids[b * max_len + i];
    const int c = counts[b * max_len + i];
    if (c > 0 && id >= 0 && id < vocab) {
      float x = Num<T>::to_f(row[id]);
      x -= (c * f);
      x -= pz;
      row[id] = Num<T>::from_f(x);
    }
  }
}

// In-place softmax in the reference's loop shape (softmax_kernels.cu:11-54): BD = min(vocab, 1024)
// threads stride the row; the exponentials are stored to the row in T and READ BACK for the sum
// (so the sum is over rounded values); the sum's butterflies are those of reduce_kernel_utils.cuh;
// the divisor gets + 1e-6 and — `logits[i] /= s_sum_val` being operator/=(T&, const T&) — is
// rounded to T before the division.
template <typename T>
__global__ void __launch_bounds__(1024) softmax_kernel(T* __restrict__ logits, int64_t vocab) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[32];
  T* row = logits + (int64_t)blockIdx.x * vocab;
  const int64_t BD = vocab < 1024 ? vocab : 1024;
  const int lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  const bool active = (int64_t)threadIdx.x < BD;
  float mx = -3.402823466e+38f;
  if (active)
    for (int64_t i = threadIdx.x; i < vocab; i += BD) mx = fmaxf(mx, Num<T>::to_f(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  float t = lane < nw ? red[lane] : -1e20f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  const float row_max = t;
  __syncthreads();  // red[] is reused
  float sum = 0.f;
  if (active)
    for (int64_t i = threadIdx.x; i < vocab; i += BD) {
      const T e = Num<T>::from_f(__expf(Num<T>::to_f(row[i]) - row_max));
      row[i] = e;
      sum += Num<T>::to_f(e);
    }
  sum = warp_sum(sum);
  if (lane == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  t = lane < nw ? red[lane] : 0.f;
  const float denom = rnd<T>(warp_sum(t) + 1e-6f);
  if (active)
    for (int64_t i = threadIdx.x; i < vocab; i += BD) row[i] = Num<T>::from_f(Num<T>::to_f(row[i]) / denom);
}


// ---------------------------------------------------------------------------------------------
// top-k / top-p filter (TopKTopPLogitsProcessor::forward, src/sampling/logits_processor.h:243-276: sort
// descending, mask positions >= k, softmax of what is left, mask positions whose exclusive cumulative
// probability exceeds p, scatter back — five library launches and a [batch, vocab] sort).  Both masks
// keep a PREFIX of the sorted order, so the result is "keep the m largest logits": this kernel finds
// m per row without sorting — a two-level radix histogram (count + probability mass per bin) over the
// 16-bit order-preserving key of the bf16 / fp16 logit — and writes -inf over everything else, in place.
//   * masses are accumulated in 2^-40 fixed point (integer adds commute: deterministic);
//   * equal logits are exact ties (same key, same mass): of a tied group that straddles the cut, the
//     ones with the lowest vocabulary index are kept (the reference's sort order among ties is
//     unspecified; its test, logits_processor_test.cpp:263-357, compares sorted values only);
//   * sums are fp32 / fixed point where the reference's are bf16 tensors (softmax and cumsum outputs
//     rounded to T): a token whose exclusive cumulative probability is within bf16 rounding of p may
//     fall on the other side (tests/test_gpu_sampling.py bounds that).
// One block of 1024 threads per row, five passes over the row (L2 resident after the first).
// ---------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 1024, TK_WARPS = 32, TK_BINS = 256;
constexpr double TK_FIX = 1099511627776.0;   // 2^40

__device__ __forceinline__ uint32_t tk_key(uint32_t u16) {   // ascending order-preserving key
  return (u16 & 0x8000u) ? (~u16 & 0xFFFFu) : (u16 | 0x8000u);
}
__device__ __forceinline__ uint32_t tk_unkey(uint32_t key) {
  return (key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu);
}
template <typename T>
__device__ __forceinline__ float tk_val(uint32_t u16) {
  const uint16_t h = (uint16_t)u16;
  return Num<T>::to_f(*reinterpret_cast<const T*>(&h));
}
__device__ __forceinline__ unsigned long long tk_mass(float x, float xmax) {
  const float e = __expf(x - xmax);
  return e >= 0.f ? (unsigned long long)((double)e * TK_FIX) : 0ull;   // NaN logits carry no mass
}

struct TkShared {
  uint32_t cnt[TK_BINS];
  unsigned long long mass[TK_BINS];
  float red[TK_WARPS];
  int red_i[TK_WARPS];
  // decisions broadcast by thread 0
  int bin_k, bin_p, need_level2_p, key_t, ties_kept, ties_total, done;
  unsigned long long m_above, z_fix, t_fix;
  uint32_t c_above;
  float xmax;
};

// The row as 16-byte vectors (8 logits) when its base and length allow, scalars otherwise: a single block
// streaming 256 KB with 2-byte loads is latency bound (~40 us per pass at vocabulary 128 256).
struct TkRow {
  const uint16_t* p;
  int n, nvec;   // nvec vectors of 8 cover [0, 8 nvec); the tail [8 nvec, n) is read as scalars
};
__device__ __forceinline__ TkRow tk_row(const uint16_t* row, int n) {
  const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
  return TkRow{row, n, vec ? n / 8 : 0};
}

// privatised histograms: [warp][bin] in dynamic shared memory, reduced into TkShared::cnt / mass.
// Level 1 (high byte: sign + 7 exponent bits — logits crowd into a handful of bins): the lanes of a warp
// that hit the same bin are found with ballots and their masses added with warp reductions, one shared-
// memory atomic per distinct bin instead of up to 32 conflicting ones.  Level 2 (low byte inside one
// high-byte bin: few lanes, spread out): plain atomics.
template <typename T, int LEVEL>
__device__ __forceinline__ void tk_add(uint32_t u, bool valid, float xmax, int bin_hi, uint32_t* wc,
                                       unsigned long long* wm) {
  const uint32_t key = tk_key(u);
  if (LEVEL == 1) {
    const int bin = valid ? (int)(key >> 8) : -1;
    const unsigned long long q = valid ? tk_mass(tk_val<T>(u), xmax) : 0ull;
    const uint32_t q_lo = (uint32_t)(q & 0xFFFFFu), q_hi = (uint32_t)(q >> 20);   // 32 x 2^20 < 2^32
    uint32_t todo = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31;
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int lbin = __shfl_sync(0xffffffffu, bin, leader);
      const bool mine = bin == lbin;
      const uint32_t peers = __ballot_sync(0xffffffffu, mine);
      const uint32_t lo = __reduce_add_sync(0xffffffffu, mine ? q_lo : 0u);
      const uint32_t hi = __reduce_add_sync(0xffffffffu, mine ? q_hi : 0u);
      if (lane == leader) {
        wc[lbin] += __popc(peers);                       // this warp's private histogram: no atomics needed
        wm[lbin] += ((unsigned long long)hi << 20) + lo;
      }
      todo &= ~peers;
    }
    __syncwarp();
  } else {
    if (valid && (int)(key >> 8) == bin_hi) {
      atomicAdd(&wc[key & 255], 1u);
      atomicAdd(&wm[key & 255], tk_mass(tk_val<T>(u), xmax));
    }
  }
}

template <typename T, int LEVEL>
__device__ __forceinline__ void tk_histogram(const TkRow& r, float xmax, int bin_hi, uint32_t* wcnt,
                                             unsigned long long* wmass, TkShared& sh) {
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < TK_WARPS * TK_BINS; i += TK_THREADS) {
    wcnt[i] = 0;
    wmass[i] = 0;
  }
  __syncthreads();
  uint32_t* wc = wcnt + warp * TK_BINS;
  unsigned long long* wm = wmass + warp * TK_BINS;
  // whole warps iterate together (the level-1 path uses warp collectives): pad the trip counts
  const int vec_iters = (r.nvec + TK_THREADS - 1) / TK_THREADS;
  // the next iteration's vector is requested before this one's is consumed (one block per row: the
  // pass is bound by the latency of its own loads otherwise)
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if ((int)threadIdx.x < r.nvec) nxt = reinterpret_cast<const uint4*>(r.p)[threadIdx.x];
  for (int it = 0; it < vec_iters; ++it) {
    const int v = it * TK_THREADS + threadIdx.x;
    const bool ok = v < r.nvec;
    const uint4 raw = nxt;
    if (v + TK_THREADS < r.nvec) nxt = reinterpret_cast<const uint4*>(r.p)[v + TK_THREADS];
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tk_add<T, LEVEL>(w[j] & 0xFFFFu, ok, xmax, bin_hi, wc, wm);
      tk_add<T, LEVEL>(w[j] >> 16, ok, xmax, bin_hi, wc, wm);
    }
  }
  const int tail0 = r.nvec * 8, tail_iters = (r.n - tail0 + TK_THREADS - 1) / TK_THREADS;
  for (int it = 0; it < tail_iters; ++it) {
    const int i = tail0 + it * TK_THREADS + threadIdx.x;
    const bool ok = i < r.n;
    tk_add<T, LEVEL>(ok ? r.p[i] : 0u, ok, xmax, bin_hi, wc, wm);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < TK_BINS; b += TK_THREADS) {
    uint32_t c = 0;
    unsigned long long m = 0;
    for (int w = 0; w < TK_WARPS; ++w) {
      c += wcnt[w * TK_BINS + b];
      m += wmass[w * TK_BINS + b];
    }
    sh.cnt[b] = c;
    sh.mass[b] = m;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(TK_THREADS) topk_topp_kernel(T* __restrict__ logits, const int64_t* __restrict__ top_k,
                                                               const float* __restrict__ top_p, int64_t vocab,
                                                               int64_t stride) {
  extern __shared__ __align__(16) uint8_t tk_dyn[];
  unsigned long long* wmass = reinterpret_cast<unsigned long long*>(tk_dyn);
  uint32_t* wcnt = reinterpret_cast<uint32_t*>(wmass + TK_WARPS * TK_BINS);
  __shared__ TkShared sh;
  pdl_wait();
  pdl_launch_dependents();
  const int n = (int)vocab;
  uint16_t* row = reinterpret_cast<uint16_t*>(logits + (int64_t)blockIdx.x * stride);
  const TkRow rr = tk_row(row, n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t k64 = top_k ? top_k[blockIdx.x] : 0;
  if (k64 <= 0 || k64 > n) k64 = n;   // <= 0 disables top-k (logits_processor.h:232-234)
  const int k = (int)k64;
  const float p = top_p ? top_p[blockIdx.x] : 2.0f;
  if (k == n && !(p < 1.0f)) return;   // nothing to filter (block-uniform)

  // ---- pass 0: row maximum ----
  float mx = -INFINITY;
#pragma unroll 4
  for (int v = threadIdx.x; v < rr.nvec; v += TK_THREADS) {
    const uint4 raw = reinterpret_cast<const uint4*>(row)[v];
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) mx = fmaxf(mx, fmaxf(tk_val<T>(w[j] & 0xFFFFu), tk_val<T>(w[j] >> 16)));
  }
  for (int i = rr.nvec * 8 + threadIdx.x; i < n; i += TK_THREADS) mx = fmaxf(mx, tk_val<T>(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sh.red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = sh.red[0];
    for (int w = 1; w < TK_WARPS; ++w) v = fmaxf(v, sh.red[w]);
    sh.xmax = v;
  }
  __syncthreads();
  const float xmax = sh.xmax;
  if (xmax == -INFINITY) return;   // an all -inf row stays as it is

  // ---- pass A: level-1 histogram (high byte of the key); the top-k boundary's bin ----
  tk_histogram<T, 1>(rr, xmax, 0, wcnt, wmass, sh);
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    unsigned long long m = 0;
    int b = TK_BINS - 1;
    for (; b > 0; --b) {
      if (c + sh.cnt[b] >= (uint32_t)k) break;
      c += sh.cnt[b];
      m += sh.mass[b];
    }
    sh.bin_k = b;
    sh.c_above = c;
    sh.m_above = m;
  }
  __syncthreads();
  const int bin_k = sh.bin_k;
  // level-1 masses of the bins above bin_k are needed again after pass B overwrites the histogram
  unsigned long long my_mass1 = 0;
  uint32_t my_cnt1 = 0;
  if (threadIdx.x < TK_BINS) {
    my_mass1 = sh.mass[threadIdx.x];
    my_cnt1 = sh.cnt[threadIdx.x];
  }
  __syncthreads();

  // ---- pass B: level-2 histogram inside bin_k: the exact k-th key, Z of the top-k set ----
  tk_histogram<T, 2>(rr, xmax, bin_k, wcnt, wmass, sh);
  if (threadIdx.x == 0) {
    uint32_t c = sh.c_above;
    unsigned long long m = sh.m_above;
    int l = TK_BINS - 1;
    for (; l > 0; --l) {
      if (c + sh.cnt[l] >= (uint32_t)k) break;
      c += sh.cnt[l];
      m += sh.mass[l];
    }
    const int key_k = (bin_k << 8) | l;
    const uint32_t need_k = (uint32_t)k - c;                      // ties of key_k inside the top k (>= 1)
    const unsigned long long q_k = sh.cnt[l] ? sh.mass[l] / sh.cnt[l] : 0;   // every tie has the same mass
    sh.z_fix = m + (unsigned long long)need_k * q_k;               // softmax denominator after the top-k mask
    sh.key_t = key_k;
    sh.ties_kept = (int)need_k;
    sh.ties_total = (int)sh.cnt[l];
    sh.t_fix = (p < 1.0f) ? (p > 0.f ? (unsigned long long)((double)p * (double)sh.z_fix) : 0ull) : ~0ull;
    sh.bin_p = -1;
    sh.need_level2_p = 0;
  }
  __syncthreads();

  // ---- top-p: the level-1 bin (above bin_k) where the exclusive cumulative mass first exceeds p Z ----
  if (p < 1.0f) {
    // serial scan over <= 256 bins by warp 0 lane 0, reading the saved level-1 masses through shared memory
    __shared__ unsigned long long m1[TK_BINS];
    __shared__ uint32_t c1[TK_BINS];
    if (threadIdx.x < TK_BINS) {
      m1[threadIdx.x] = my_mass1;
      c1[threadIdx.x] = my_cnt1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long M = 0;
      uint32_t C = 0;
      int found = -1;
      for (int b = TK_BINS - 1; b > bin_k; --b) {
        if (M + m1[b] > sh.t_fix) { found = b; break; }
        M += m1[b];
        C += c1[b];
      }
      if (found >= 0) {          // the cut lies inside a bin above bin_k: needs that bin's level-2 histogram
        sh.bin_p = found;
        sh.need_level2_p = 1;
        sh.m_above = M;
        sh.c_above = C;
      } else {                   // inside bin_k (its top-k part), or not at all
        sh.bin_p = bin_k;
        // sh.m_above / c_above still describe the bins above bin_k (== M, C)
      }
    }
    __syncthreads();
    if (sh.need_level2_p) tk_histogram<T, 2>(rr, xmax, sh.bin_p, wcnt, wmass, sh);   // block-uniform branch
    if (threadIdx.x == 0) {
      unsigned long long M = sh.m_above;
      const int key_k = sh.key_t;
      const int in_k_bin = sh.bin_p == bin_k;
      for (int l = TK_BINS - 1; l >= 0; --l) {
        const int key = (sh.bin_p << 8) | l;
        if (in_k_bin && key < key_k) break;                                   // beyond the top-k set
        uint32_t c = sh.cnt[l];
        if (c == 0) continue;
        const unsigned long long q = sh.mass[l] / c;
        if (in_k_bin && key == key_k) c = (uint32_t)sh.ties_kept;             // only need_k of these ties are in the top k
        if (M + (unsigned long long)c * q > sh.t_fix) {
          // positions t = 0.. of this tied group have exclusive mass M + t q: kept while <= p Z
          unsigned long long kept = q ? (sh.t_fix - M) / q + 1 : c;
          if (kept > c) kept = c;
          sh.key_t = key;
          sh.ties_kept = (int)kept;
          sh.ties_total = (int)sh.cnt[l];
          break;
        }
        M += (unsigned long long)c * q;
        if (in_k_bin && key == key_k) break;
      }
    }
    __syncthreads();
  }

  // ---- pass D: keep key > key_t, and the first ties_kept (by index) of key == key_t ----
  const int key_t = sh.key_t, ties_kept = sh.ties_kept;
  const bool rank_ties = ties_kept < sh.ties_total;
  const uint32_t ninf = std::is_same<T, __nv_bfloat16>::value ? 0xFF80u : 0xFC00u;
  int base = 0;   // ties seen in earlier chunks (block-uniform)
  // chunks of 8 elements per thread (vector part) then 1 per thread (tail): index order = thread order
  const int vec_iters = (rr.nvec + TK_THREADS - 1) / TK_THREADS;
  const int tail0 = rr.nvec * 8, tail_iters = (n - tail0 + TK_THREADS - 1) / TK_THREADS;
  for (int it = 0; it < vec_iters + tail_iters; ++it) {
    const bool vec = it < vec_iters;
    const int v = it * TK_THREADS + threadIdx.x;                         // vector index (vec part)
    const int i1 = tail0 + (it - vec_iters) * TK_THREADS + threadIdx.x;   // element index (tail part)
    const bool ok = vec ? v < rr.nvec : i1 < n;
    uint32_t e[8];
    int cnt = vec ? 8 : 1;
    if (ok && vec) {
      const uint4 raw = reinterpret_cast<const uint4*>(row)[v];
      e[0] = raw.x & 0xFFFFu; e[1] = raw.x >> 16; e[2] = raw.y & 0xFFFFu; e[3] = raw.y >> 16;
      e[4] = raw.z & 0xFFFFu; e[5] = raw.z >> 16; e[6] = raw.w & 0xFFFFu; e[7] = raw.w >> 16;
    } else if (ok) {
      e[0] = row[i1];
    }
    if (!ok) cnt = 0;
    int my_ties = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < cnt && (int)tk_key(e[j]) == key_t) ++my_ties;
    int before = 0;
    if (rank_ties) {   // block-uniform
      // exclusive prefix of my_ties over the block in thread order
      int incl = my_ties;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) sh.red_i[warp] = incl;
      __syncthreads();
      before = base + incl - my_ties;
      int total = 0;
      for (int w = 0; w < TK_WARPS; ++w) {
        if (w < warp) before += sh.red_i[w];
        total += sh.red_i[w];
      }
      base += total;
      __syncthreads();
    }
    bool changed = false;
    int seen = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= cnt) continue;
      const int key = (int)tk_key(e[j]);
      bool keep = key > key_t;
      if (key == key_t) {
        keep = !rank_ties || (before + seen) < ties_kept;
        ++seen;
      }
      if (!keep) {
        e[j] = ninf;
        changed = true;
      }
    }
    if (changed) {
      if (vec) {
        reinterpret_cast<uint4*>(row)[v] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16),
                                                      e[6] | (e[7] << 16));
      } else {
        row[i1] = (uint16_t)e[0];
      }
    }
  }
}

}  // namespace b200

using namespace b200;

#define SAMPLING_DISPATCH(dtype, ...)                                  \
  switch (dtype) {                                                     \
    case B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }   \
    case B200_FP16: { using T = __half; __VA_ARGS__; break; }          \
    case B200_FP32: { using T = float; __VA_ARGS__; break; }           \
    default: return set_error(B200_ERR_INVALID_ARG, "bad dtype %d", dtype); \
  }

extern "C" {

int b200_apply_temperature(void* logits, const void* temperatures, int64_t batch, int64_t vocab,
                           int dtype, b200_stream_t stream) {
  if (batch == 0) return B200_OK;
  B200_CHECK_ARG(logits && temperatures && batch > 0 && vocab > 0, "apply_temperature: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (batch * vocab + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  SAMPLING_DISPATCH(dtype, B200_PDL_LAUNCH("apply_temperature", temperature_kernel<T>, (unsigned)blocks, 256, 0, st,
                                           static_cast<T*>(logits), static_cast<const T*>(temperatures), batch, vocab));
  return B200_OK;
}

int b200_apply_repetition_penalty(void* logits, const int64_t* token_ids, const int32_t* token_ids_lens,
                                  const void* penalties, int64_t batch, int64_t vocab, int64_t max_len,
                                  int dtype, b200_stream_t stream) {
  if (batch == 0 || max_len == 0) return B200_OK;
  B200_CHECK_ARG(logits && token_ids && token_ids_lens && penalties && batch > 0 && vocab > 0 && max_len > 0,
                 "apply_repetition_penalty: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  SAMPLING_DISPATCH(dtype, B200_PDL_LAUNCH("apply_repetition_penalty", repetition_penalty_kernel<T>, (unsigned)batch,
                                           256, 0, st, static_cast<T*>(logits), token_ids, token_ids_lens,
                                           static_cast<const T*>(penalties), max_len, vocab));
  return B200_OK;
}

int b200_apply_frequency_presence_penalty(void* logits, const int64_t* token_ids, const int32_t* token_counts,
                                          const int32_t* token_ids_lens, const void* frequency_penalties,
                                          const void* presence_penalties, int64_t batch, int64_t vocab,
                                          int64_t max_len, int dtype, b200_stream_t stream) {
  if (batch == 0 || max_len == 0) return B200_OK;
  B200_CHECK_ARG(logits && token_ids && token_counts && token_ids_lens && frequency_penalties &&
                     presence_penalties && batch > 0 && vocab > 0 && max_len > 0,
                 "apply_frequency_presence_penalty: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  SAMPLING_DISPATCH(dtype, B200_PDL_LAUNCH("apply_frequency_presence_penalty", frequency_presence_penalty_kernel<T>,
                                           (unsigned)batch, 256, 0, st, static_cast<T*>(logits), token_ids,
                                           token_counts, token_ids_lens, static_cast<const T*>(frequency_penalties),
                                           static_cast<const T*>(presence_penalties), max_len, vocab));
  return B200_OK;
}

int b200_softmax(void* logits, int64_t batch, int64_t vocab, int dtype, b200_stream_t stream) {
  if (batch == 0) return B200_OK;
  B200_CHECK_ARG(logits && batch > 0 && vocab > 0, "softmax: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  const int threads = (int)(vocab < 1024 ? ((vocab + 31) / 32) * 32 : 1024);
  SAMPLING_DISPATCH(dtype, B200_PDL_LAUNCH("softmax", softmax_kernel<T>, (unsigned)batch, threads, 0, st,
                                           static_cast<T*>(logits), vocab));
  return B200_OK;
}

int b200_topk_topp_filter(void* logits, const int64_t* top_k, const float* top_p, int64_t batch, int64_t vocab,
                          int64_t stride, int dtype, b200_stream_t stream) {
  if (batch == 0 || (top_k == nullptr && top_p == nullptr)) return B200_OK;
  B200_CHECK_ARG(logits && batch > 0 && vocab > 0 && vocab < (1ll << 31) && stride >= vocab,
                 "topk_topp_filter: bad arguments");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "topk_topp_filter: bf16 / fp16 logits only");
  auto st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)TK_WARPS * TK_BINS * (sizeof(unsigned long long) + sizeof(uint32_t));
  if (dtype == B200_BF16) {
    auto kern = topk_topp_kernel<__nv_bfloat16>;
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_PDL_LAUNCH("topk_topp_filter", kern, (unsigned)batch, TK_THREADS, smem, st,
                    static_cast<__nv_bfloat16*>(logits), top_k, top_p, vocab, stride);
  } else {
    auto kern = topk_topp_kernel<__half>;
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_PDL_LAUNCH("topk_topp_filter", kern, (unsigned)batch, TK_THREADS, smem, st, static_cast<__half*>(logits),
                    top_k, top_p, vocab, stride);
  }
  return B200_OK;
}

}  // extern "C"
