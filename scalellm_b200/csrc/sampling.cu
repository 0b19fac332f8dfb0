// sampling.cu — the sampling tail's logits processors on the device:
// temperature, repetition / frequency / presence penalties and the in-place softmax, with the
// semantics (and rounding points) of the reference's kernels
//   src/kernels/sampling/penalty_kernels.cu:9-33,52-75,107-140   src/kernels/sampling/softmax_kernels.cu:11-54
// restated, never copied.  Greedy selection is b200_argmax / b200_ar_argmax (elementwise.cu,
// allreduce.cu).  All in place on logits [batch, vocab] (contiguous), one launch each, capturable.
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

}  // extern "C"
