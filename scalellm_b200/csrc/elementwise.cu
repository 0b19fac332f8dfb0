// elementwise.cu — RMSNorm, RoPE, KV-slot write/gather, SiLU(*mul) for sm_100a.
//
// All of these are HBM/latency-bound byte movers.  Design
// rules applied: 128-bit coalesced global accesses, the row held in registers so
// the input is read exactly once, fp32 math with the reference's per-op rounding
// reproduced exactly, and launch shapes that depend only on host scalars (CUDA
// graph capturable).
//
// Reference semantics restated (never copied):
//   RMSNorm         src/kernels/layernorm_kernels.cu:15-41,125-155
//   RoPE            src/kernels/pos_embedding_kernels.cu:10-82
//   KV write        src/kernels/kv_cache_kernels.cu:9-41
//   SiLU / SiLU*mul src/kernels/activation_kernels.cu:44-50,53-95

#include "common.cuh"

namespace b200 {

// ===========================================================================
// RMSNorm
// ===========================================================================
// Vector path: n % VEC == 0, one CTA per row, each thread owns up to MAXV
// 16-byte vectors of the row in registers (row read once).
template <typename T, int THREADS, int MAXV, bool RESIDUAL>
__global__ void __launch_bounds__(THREADS) rms_norm_vec_kernel(T* __restrict__ out,
                                                               T* __restrict__ residual,
                                                               const T* __restrict__ in,
                                                               const T* __restrict__ weight,
                                                               float eps, int n) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[32];
  extern __shared__ float sq[];  // [n] fp32: the row, for the reference-ordered sum of squares
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  const int nvec = n / VEC;
  const T* in_row = in + row * n;
  T* res_row = RESIDUAL ? residual + row * n : nullptr;
  T* out_row = out + row * n;

  float x[MAXV][VEC];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < nvec) {
      uint4 raw = ld_nc_v4(in_row + v * VEC);
      const T* e = reinterpret_cast<const T*>(&raw);
      if constexpr (RESIDUAL) {
        uint4 rraw = ld_v4(res_row + v * VEC);
        const T* r = reinterpret_cast<const T*>(&rraw);
        uint4 sraw;
        T* s = reinterpret_cast<T*>(&sraw);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          // x = float(r) + float(in): variance uses the UNROUNDED fp32 sum, the
          // second pass re-reads the rounded residual (layernorm_kernels.cu:137-153)
          const float f = Num<T>::to_f(r[i]) + Num<T>::to_f(e[i]);
          sq[v * VEC + i] = f;
          s[i] = Num<T>::from_f(f);
          x[j][i] = Num<T>::to_f(s[i]);
        }
        st_v4(res_row + v * VEC, sraw);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          x[j][i] = Num<T>::to_f(e[i]);
          sq[v * VEC + i] = x[j][i];
        }
      }
    }
  }
  const float total = row_sumsq_ref_order<THREADS>(sq, red, n);
  const float rstd = rsqrtf(total / n + eps);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < nvec) {
      uint4 wraw = ld_v4(weight + v * VEC);
      const T* w = reinterpret_cast<const T*>(&wraw);
      uint4 oraw;
      T* o = reinterpret_cast<T*>(&oraw);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        // (T)(x * rstd) * w  — two roundings (layernorm_kernels.cu:39)
        const float y = rnd<T>(x[j][i] * rstd);
        o[i] = Num<T>::from_f(y * Num<T>::to_f(w[i]));
      }
      st_v4(out_row + v * VEC, oraw);
    }
  }
}

// Scalar fallback: any n (the reference's own test uses n = 1038).
template <typename T, bool RESIDUAL>
__global__ void __launch_bounds__(1024) rms_norm_scalar_kernel(T* __restrict__ out,
                                                               T* __restrict__ residual,
                                                               const T* __restrict__ in,
                                                               const T* __restrict__ weight,
                                                               float eps, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  // the reference's own loop shape: BD = min(n, 1024) threads stride the row, one FFMA per element
  // (threads >= BD of our warp-rounded block idle), then its two butterflies (reduce_kernel_utils.cuh:41-64)
  const int64_t BD = n < 1024 ? n : 1024;
  float ss = 0.f;
  if ((int64_t)threadIdx.x < BD) {
    for (int64_t i = threadIdx.x; i < n; i += BD) {
      float f = Num<T>::to_f(in[row * n + i]);
      if constexpr (RESIDUAL) {
        f = Num<T>::to_f(residual[row * n + i]) + f;
        residual[row * n + i] = Num<T>::from_f(f);
      }
      ss = fmaf(f, f, ss);
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float t = (threadIdx.x & 31) < ((blockDim.x + 31) >> 5) ? red[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  const float rstd = rsqrtf(t / n + eps);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float f = RESIDUAL ? Num<T>::to_f(residual[row * n + i]) : Num<T>::to_f(in[row * n + i]);
    const float y = rnd<T>(f * rstd);
    out[row * n + i] = Num<T>::from_f(y * Num<T>::to_f(weight[i]));
  }
}

// Gemma RMSNorm and LayerNorm (layernorm_kernels.cu:66-95,185-230): adjacent to the Llama path (the
// reference's Gemma / GPT-2 / Phi models link them from the same :kernels target).  Kept in the
// reference's own loop shape — BD = min(n, 1024) threads stride the row, FFMA accumulation, the two
// butterflies — so the results are bit-identical; they are not on the benchmarked path.
// gemma: out = (T)(x * rstd * (1.0 + w)) with the (1.0 + w) factor in DOUBLE, as written there.
template <typename T>
__global__ void __launch_bounds__(1024) gemma_rms_norm_kernel(T* __restrict__ out,
                                                              const T* __restrict__ in,
                                                              const T* __restrict__ weight,
                                                              float eps, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int64_t BD = n < 1024 ? n : 1024;
  float ss = 0.f;
  if ((int64_t)threadIdx.x < BD)
    for (int64_t i = threadIdx.x; i < n; i += BD) {
      const float f = Num<T>::to_f(in[row * n + i]);
      ss = fmaf(f, f, ss);
    }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float t = (threadIdx.x & 31) < ((blockDim.x + 31) >> 5) ? red[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  const float rstd = rsqrtf(t / n + eps);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = Num<T>::to_f(in[row * n + i]);
    const float w = Num<T>::to_f(weight[i]);
    out[row * n + i] = Num<T>::from_f((float)(x * rstd * (1.0 + w)));
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) layer_norm_kernel(T* __restrict__ out,
                                                          const T* __restrict__ in,
                                                          const T* __restrict__ weight,
                                                          const T* __restrict__ bias, float eps,
                                                          int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[32];
  __shared__ float s_mean;
  const int64_t row = blockIdx.x;
  const int64_t BD = n < 1024 ? n : 1024;
  const int lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  float m = 0.f;
  if ((int64_t)threadIdx.x < BD)
    for (int64_t i = threadIdx.x; i < n; i += BD) m += Num<T>::to_f(in[row * n + i]);
  m = warp_sum(m);
  if (lane == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.f;
  t = warp_sum(t);
  if (threadIdx.x == 0) s_mean = t / n;
  __syncthreads();
  const float mean = s_mean;
  float var = 0.f;
  if ((int64_t)threadIdx.x < BD)
    for (int64_t i = threadIdx.x; i < n; i += BD) {
      const float x = Num<T>::to_f(in[row * n + i]) - mean;
      var = fmaf(x, x, var);
    }
  var = warp_sum(var);
  __syncthreads();  // red[] is reused
  if (lane == 0) red[threadIdx.x >> 5] = var;
  __syncthreads();
  t = lane < nw ? red[lane] : 0.f;
  t = warp_sum(t);
  const float rstd = rsqrtf(t / n + eps);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float o = (Num<T>::to_f(in[row * n + i]) - mean) * rstd * Num<T>::to_f(weight[i]);
    if (bias != nullptr) o += Num<T>::to_f(bias[i]);
    out[row * n + i] = Num<T>::from_f(o);
  }
}

template <typename T, bool RESIDUAL>
static int launch_rms_norm(void* out, void* residual, const void* in, const void* weight,
                           int64_t rows, int64_t n, float eps, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (n % VEC == 0) && is_aligned(out, 16) && is_aligned(in, 16) &&
                      is_aligned(weight, 16) && (!RESIDUAL || is_aligned(residual, 16));
  const int64_t nvec = n / VEC;
  T* o = static_cast<T*>(out);
  T* r = static_cast<T*>(residual);
  const T* i = static_cast<const T*>(in);
  const T* w = static_cast<const T*>(weight);
  dim3 grid(static_cast<unsigned>(rows));
  const size_t sm = (size_t)n * 4;  // the row in fp32 (reference-ordered sum of squares)
  if (vec_ok && nvec <= 128) {
    B200_PDL_LAUNCH("rms_norm", (rms_norm_vec_kernel<T, 128, 1, RESIDUAL>), grid, 128, sm, st, o, r, i, w, eps, (int)n);
  } else if (vec_ok && nvec <= 256) {
    B200_PDL_LAUNCH("rms_norm", (rms_norm_vec_kernel<T, 256, 1, RESIDUAL>), grid, 256, sm, st, o, r, i, w, eps, (int)n);
  } else if (vec_ok && nvec <= 512) {
    B200_PDL_LAUNCH("rms_norm", (rms_norm_vec_kernel<T, 512, 1, RESIDUAL>), grid, 512, sm, st, o, r, i, w, eps, (int)n);
  } else if (vec_ok && nvec <= 1024) {
    B200_PDL_LAUNCH("rms_norm", (rms_norm_vec_kernel<T, 512, 2, RESIDUAL>), grid, 512, sm, st, o, r, i, w, eps, (int)n);
  } else if (vec_ok && sm <= 48 * 1024) {   // larger rows: the strided kernel below (same order, no staging)
    B200_PDL_LAUNCH("rms_norm", (rms_norm_vec_kernel<T, 1024, 4, RESIDUAL>), grid, 1024, sm, st, o, r, i, w, eps, (int)n);
  } else {
    const int threads = (int)((n < 1024 ? ((n + 31) / 32) * 32 : 1024));
    B200_PDL_LAUNCH("rms_norm", (rms_norm_scalar_kernel<T, RESIDUAL>), grid, threads, 0, st, o, r, i, w, eps, n);
  }
  return B200_OK;
}


// Residual RMSNorm whose `in` operand arrives as the W4A16 GEMM's stream-K partials in fp32:
// x = T(sum over the tile's contributor slots of P[slot][row][:]) — the single rounding a GEMM
// epilogue would have done — then exactly rms_norm_residual.  One CTA per row, row in registers.
template <typename T, int THREADS, int MAXV>
__global__ void __launch_bounds__(THREADS) rms_norm_residual_splitk_kernel(
    T* __restrict__ out, T* __restrict__ residual, const float* __restrict__ partials, W4Plan plan,
    int64_t split_stride, const T* __restrict__ weight, float eps, int n) {
  constexpr int VEC = 16 / sizeof(T);
  static_assert(VEC == 8, "16-bit element types only");
  __shared__ float red[32];
  extern __shared__ float sq[];  // [n] fp32: the row, for the reference-ordered sum of squares
  const int64_t row = blockIdx.x;
  const int nvec = n / VEC;
  T* res_row = residual + row * n;
  T* out_row = out + row * n;
  const float* p_row = partials + row * n;
  // Before griddepcontrol.wait: everything that does not depend on the producing GEMM — the
  // residual row (last written by the previous norm kernel, several launches back), the norm
  // weights and the stream-K contributor counts; their latency overlaps the GEMM's tail.
  uint4 rraw[MAXV], wraw[MAXV];
  int cnt[MAXV];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < nvec) {
      rraw[j] = ld_v4(res_row + v * VEC);
      wraw[j] = ld_v4(weight + v * VEC);
      cnt[j] = w4_contrib_col(plan, v * VEC);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  float x[MAXV][VEC];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < nvec) {
      float a[VEC];
      w4_sum_partials8(a, p_row + v * VEC, split_stride, cnt[j]);
      const T* r = reinterpret_cast<const T*>(&rraw[j]);
      uint4 sraw;
      T* sv = reinterpret_cast<T*>(&sraw);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float gemm_out = rnd<T>(a[i]);  // what the GEMM would have stored
        const float f = Num<T>::to_f(r[i]) + gemm_out;
        sq[v * VEC + i] = f;
        sv[i] = Num<T>::from_f(f);
        x[j][i] = Num<T>::to_f(sv[i]);
      }
      st_v4(res_row + v * VEC, sraw);
    }
  }
  const float total = row_sumsq_ref_order<THREADS>(sq, red, n);
  const float rstd = rsqrtf(total / n + eps);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < nvec) {
      const T* w = reinterpret_cast<const T*>(&wraw[j]);
      uint4 oraw;
      T* o = reinterpret_cast<T*>(&oraw);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float y = rnd<T>(x[j][i] * rstd);
        o[i] = Num<T>::from_f(y * Num<T>::to_f(w[i]));
      }
      st_v4(out_row + v * VEC, oraw);
    }
  }
}

template <typename T>
static int launch_rms_norm_splitk(void* out, void* residual, const float* partials, W4Plan S,
                                  int64_t split_stride, const void* weight, int64_t rows,
                                  int64_t n, float eps, cudaStream_t st) {
  T* o = static_cast<T*>(out);
  T* r = static_cast<T*>(residual);
  const T* w = static_cast<const T*>(weight);
  const int64_t nvec = n / 8;
  dim3 grid(static_cast<unsigned>(rows));
  const size_t sm = (size_t)n * 4;  // the row in fp32 (reference-ordered sum of squares)
  if (sm > 48 * 1024) {             // n <= 32768 (checked by the caller): opt in to the larger carve-out
    B200_CUDA_OK(cudaFuncSetAttribute(rms_norm_residual_splitk_kernel<T, 1024, 4>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  }
  if (nvec <= 256)
    B200_PDL_LAUNCH_L(1, "rms_norm_residual_splitk", (rms_norm_residual_splitk_kernel<T, 256, 1>), grid, 256, sm, st, o, r, partials, S, split_stride, w, eps, (int)n);
  else if (nvec <= 512)
    B200_PDL_LAUNCH_L(1, "rms_norm_residual_splitk", (rms_norm_residual_splitk_kernel<T, 512, 1>), grid, 512, sm, st, o, r, partials, S, split_stride, w, eps, (int)n);
  else if (nvec <= 1024)
    B200_PDL_LAUNCH_L(1, "rms_norm_residual_splitk", (rms_norm_residual_splitk_kernel<T, 512, 2>), grid, 512, sm, st, o, r, partials, S, split_stride, w, eps, (int)n);
  else
    B200_PDL_LAUNCH_L(1, "rms_norm_residual_splitk", (rms_norm_residual_splitk_kernel<T, 1024, 4>), grid, 1024, sm, st, o, r, partials, S, split_stride, w, eps, (int)n);
  return B200_OK;
}

// ===========================================================================
// RoPE (+ optional fused KV-slot write)
// ===========================================================================
// One CTA per token.  A work item is one 16-byte vector of the x half paired
// with the matching vector of the y half (non-interleaved), or one vector of
// interleaved (x,y) pairs.  Every multiply and add is rounded to T.
template <typename T>
__device__ __forceinline__ void rope_pair(float x, float y, float c, float s, T& ox, T& oy) {
  // x' = x*c - y*s ; y' = x*s + y*c   with per-op rounding to T
  const float xc = rnd<T>(x * c), ys = rnd<T>(y * s);
  const float xs = rnd<T>(x * s), yc = rnd<T>(y * c);
  ox = Num<T>::from_f(xc - ys);
  oy = Num<T>::from_f(xs + yc);
}

// The qkv GEMM's stream-K partials as the source of a token's [q | k | v] row (FROM_PARTIALS).
struct RopePartials {
  const float* data;     // [slots][n_tokens][row_n] fp32, row_n = (n_heads + 2 n_kv_heads) head_dim
  int64_t slot_stride;
  int row_n;
  W4Plan plan;
};

template <typename T, bool FUSE_KV, bool FROM_PARTIALS>
__global__ void __launch_bounds__(256) rope_vec_kernel(
    T* __restrict__ q, T* __restrict__ k, const T* __restrict__ v,
    const int32_t* __restrict__ positions, const T* __restrict__ cos_sin,
    const int32_t* __restrict__ slot_ids, T* __restrict__ k_cache, T* __restrict__ v_cache,
    int n_heads, int n_kv_heads, int head_dim, int rotary_dim, int64_t q_stride,
    int64_t k_stride, int64_t v_stride, bool interleaved, RopePartials parts) {
  constexpr int VEC = 16 / sizeof(T);
  const int64_t tok = blockIdx.x;
  // positions / slot ids / the cos|sin table are step inputs and model constants, older than the
  // predecessor kernel: fetch them before griddepcontrol.wait so their latency overlaps its tail
  const int half = rotary_dim / 2;
  const T* cs = cos_sin + static_cast<int64_t>(positions[tok]) * rotary_dim;
  const int64_t slot = FUSE_KV ? static_cast<int64_t>(slot_ids[tok]) : 0;
  extern __shared__ __align__(16) uint8_t rope_smem[];  // FROM_PARTIALS: the token's [q|k|v] row in T
  T* row_s = reinterpret_cast<T*>(rope_smem);
  pdl_wait();
  pdl_launch_dependents();
  if constexpr (FROM_PARTIALS) {
    // Materialise this token's qkv row first: x = T(sum of the tile's partial slots), the one
    // rounding the GEMM epilogue would have done; q, k, v are views of that row (host-checked).
    // Three vectors per thread per round so that all their partial loads are in flight together;
    // the row also stays in shared memory for the rotation below (no global re-read).
    static_assert(sizeof(T) == 2, "partials input: 16-bit element types");
    T* row = q + tok * q_stride;
    const float* prow = parts.data + tok * parts.row_n;
    const int nv = parts.row_n / 8;
    for (int v0 = threadIdx.x; v0 < nv; v0 += 3 * blockDim.x) {
      float a[3][8];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int vv = v0 + u * blockDim.x;
        if (vv < nv)
          w4_sum_partials8(a[u], prow + vv * 8, parts.slot_stride, w4_contrib_col(parts.plan, vv * 8));
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int vv = v0 + u * blockDim.x;
        if (vv < nv) {
          uint4 o;
          T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
          for (int i = 0; i < 8; ++i) oe[i] = Num<T>::from_f(a[u][i]);
          st_v4(row + vv * 8, o);
          st_v4(row_s + vv * 8, o);
        }
      }
    }
    __syncthreads();  // the row is re-read below by other threads of this CTA
  }
  // source of the un-rotated values: the row in shared memory (offsets relative to q's row start:
  // q, k, v are views of one [q|k|v] row) or global memory
  const T* q_src = FROM_PARTIALS ? row_s : q + tok * q_stride;
  const T* k_src = FROM_PARTIALS ? row_s + (k - q) : k + tok * k_stride;
  const T* v_src = FROM_PARTIALS ? row_s + (v - q) : v + tok * v_stride;
  const T* cosp = cs;
  const T* sinp = cs + half;
  T* kc_row = FUSE_KV ? k_cache + slot * n_kv_heads * head_dim : nullptr;
  T* vc_row = FUSE_KV ? v_cache + slot * n_kv_heads * head_dim : nullptr;

  // items per head: non-interleaved -> half/VEC vector pairs; interleaved -> rotary_dim/VEC vectors
  const int items_per_head = interleaved ? rotary_dim / VEC : half / VEC;
  const int total_heads = n_heads + n_kv_heads;
  for (int it = threadIdx.x; it < total_heads * items_per_head; it += blockDim.x) {
    const int h = it / items_per_head;
    const int j = it % items_per_head;
    const bool is_k = h >= n_heads;
    T* base = is_k ? k + tok * k_stride + static_cast<int64_t>(h - n_heads) * head_dim
                   : q + tok * q_stride + static_cast<int64_t>(h) * head_dim;
    const T* src = is_k ? k_src + static_cast<int64_t>(h - n_heads) * head_dim
                        : q_src + static_cast<int64_t>(h) * head_dim;
    T* cdst = (FUSE_KV && is_k) ? kc_row + static_cast<int64_t>(h - n_heads) * head_dim : nullptr;
    if (!interleaved) {
      uint4 xr = ld_v4(src + j * VEC), yr = ld_v4(src + half + j * VEC);
      uint4 cr = ld_v4(cosp + j * VEC), sr = ld_v4(sinp + j * VEC);
      const T* x = reinterpret_cast<const T*>(&xr);
      const T* y = reinterpret_cast<const T*>(&yr);
      const T* c = reinterpret_cast<const T*>(&cr);
      const T* s = reinterpret_cast<const T*>(&sr);
      uint4 oxr, oyr;
      T* ox = reinterpret_cast<T*>(&oxr);
      T* oy = reinterpret_cast<T*>(&oyr);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        rope_pair<T>(Num<T>::to_f(x[i]), Num<T>::to_f(y[i]), Num<T>::to_f(c[i]),
                     Num<T>::to_f(s[i]), ox[i], oy[i]);
      st_v4(base + j * VEC, oxr);
      st_v4(base + half + j * VEC, oyr);
      if (cdst) {
        st_v4(cdst + j * VEC, oxr);
        st_v4(cdst + half + j * VEC, oyr);
      }
    } else {
      uint4 pr = ld_v4(src + j * VEC);
      const T* p = reinterpret_cast<const T*>(&pr);
      uint4 outr;
      T* o = reinterpret_cast<T*>(&outr);
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        const int r = j * (VEC / 2) + i;
        rope_pair<T>(Num<T>::to_f(p[2 * i]), Num<T>::to_f(p[2 * i + 1]), Num<T>::to_f(cosp[r]),
                     Num<T>::to_f(sinp[r]), o[2 * i], o[2 * i + 1]);
      }
      st_v4(base + j * VEC, outr);
      if (cdst) st_v4(cdst + j * VEC, outr);
    }
  }
  if constexpr (FUSE_KV) {
    // un-rotated tail of K (rotary_dim < head_dim) and the whole of V
    const int tail_vecs = (head_dim - rotary_dim) / VEC;
    for (int it = threadIdx.x; it < n_kv_heads * tail_vecs; it += blockDim.x) {
      const int h = it / tail_vecs, j = it % tail_vecs;
      const int64_t off = static_cast<int64_t>(h) * head_dim + rotary_dim + j * VEC;
      st_v4(kc_row + off, ld_v4(k_src + off));
    }
    const int v_vecs = n_kv_heads * head_dim / VEC;
    for (int it = threadIdx.x; it < v_vecs; it += blockDim.x)
      st_v4(vc_row + static_cast<int64_t>(it) * VEC,
            FROM_PARTIALS ? ld_v4(v_src + it * VEC)               // this CTA's shared-memory row
                          : ld_nc_v4(v_src + it * VEC));
  }
}

// Scalar fallback (any rotary_dim / alignment), optional fused KV write.
template <typename T, bool FUSE_KV>
__global__ void __launch_bounds__(256) rope_scalar_kernel(
    T* __restrict__ q, T* __restrict__ k, const T* __restrict__ v,
    const int32_t* __restrict__ positions, const T* __restrict__ cos_sin,
    const int32_t* __restrict__ slot_ids, T* __restrict__ k_cache, T* __restrict__ v_cache,
    int n_heads, int n_kv_heads, int head_dim, int rotary_dim, int64_t q_stride,
    int64_t k_stride, int64_t v_stride, bool interleaved) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t tok = blockIdx.x;
  const int half = rotary_dim / 2;
  const T* cs = cos_sin + static_cast<int64_t>(positions[tok]) * rotary_dim;
  const int total = (n_heads + n_kv_heads) * half;
  for (int it = threadIdx.x; it < total; it += blockDim.x) {
    const int h = it / half, r = it % half;
    const bool is_k = h >= n_heads;
    T* base = is_k ? k + tok * k_stride + static_cast<int64_t>(h - n_heads) * head_dim
                   : q + tok * q_stride + static_cast<int64_t>(h) * head_dim;
    const int xi = interleaved ? 2 * r : r;
    const int yi = interleaved ? 2 * r + 1 : r + half;
    T ox, oy;
    rope_pair<T>(Num<T>::to_f(base[xi]), Num<T>::to_f(base[yi]), Num<T>::to_f(cs[r]),
                 Num<T>::to_f(cs[half + r]), ox, oy);
    base[xi] = ox;
    base[yi] = oy;
  }
  if constexpr (FUSE_KV) {
    __syncthreads();  // rotated K of this token is complete (same CTA wrote it)
    const int64_t slot = slot_ids[tok];
    const int nkv = n_kv_heads * head_dim;
    for (int it = threadIdx.x; it < nkv; it += blockDim.x) {
      k_cache[slot * nkv + it] = k[tok * k_stride + it];
      v_cache[slot * nkv + it] = v[tok * v_stride + it];
    }
  }
}

template <typename T, bool FUSE_KV>
static int launch_rope(void* q, void* k, const void* v, const int32_t* positions,
                       const void* cos_sin, const int32_t* slot_ids, void* k_cache, void* v_cache,
                       int64_t n_tokens, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim,
                       int64_t rotary_dim, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                       int interleaved, cudaStream_t st, const RopePartials* parts = nullptr) {
  constexpr int VEC = 16 / sizeof(T);
  if (n_tokens == 0) return B200_OK;
  const int64_t half = rotary_dim / 2;
  bool vec_ok = is_aligned(q, 16) && is_aligned(k, 16) && is_aligned(cos_sin, 16) &&
                head_dim % VEC == 0 && q_stride % VEC == 0 && k_stride % VEC == 0 &&
                rotary_dim % VEC == 0 && (interleaved ? true : half % VEC == 0);
  if (FUSE_KV)
    vec_ok = vec_ok && is_aligned(v, 16) && is_aligned(k_cache, 16) && is_aligned(v_cache, 16) &&
             v_stride % VEC == 0;
  dim3 grid(static_cast<unsigned>(n_tokens));
  auto* qq = static_cast<T*>(q);
  auto* kk = static_cast<T*>(k);
  auto* vv = static_cast<const T*>(v);
  auto* cs = static_cast<const T*>(cos_sin);
  auto* kc = static_cast<T*>(k_cache);
  auto* vc = static_cast<T*>(v_cache);
  if constexpr (FUSE_KV && sizeof(T) == 2) {
    if (parts) {
      if (!vec_ok)
        return set_error(B200_ERR_UNSUPPORTED, "rope_kv_write_splitk: needs the vectorised layout");
      const size_t row_bytes = (size_t)parts->row_n * sizeof(T);   // the [q|k|v] row in shared memory
      if (row_bytes > 48 * 1024)
        return set_error(B200_ERR_UNSUPPORTED, "rope_kv_write_splitk: qkv row of %zu bytes exceeds 48 KB", row_bytes);
      B200_PDL_LAUNCH_L(1, "rope", (rope_vec_kernel<T, true, true>), grid, 256, row_bytes, st, qq, kk, vv,
                      positions, cs, slot_ids, kc, vc, (int)n_heads, (int)n_kv_heads,
                      (int)head_dim, (int)rotary_dim, q_stride, k_stride, v_stride,
                      interleaved != 0, *parts);
      return B200_OK;
    }
  }
  if (vec_ok) {
    B200_PDL_LAUNCH("rope", (rope_vec_kernel<T, FUSE_KV, false>), grid, 256, 0, st, qq, kk, vv,
                    positions, cs, slot_ids, kc, vc, (int)n_heads, (int)n_kv_heads, (int)head_dim,
                    (int)rotary_dim, q_stride, k_stride, v_stride, interleaved != 0,
                    RopePartials{});
  } else {
    B200_PDL_LAUNCH("rope", (rope_scalar_kernel<T, FUSE_KV>), grid, 256, 0, st, qq, kk, vv,
                    positions, cs, slot_ids, kc, vc, (int)n_heads, (int)n_kv_heads, (int)head_dim,
                    (int)rotary_dim, q_stride, k_stride, v_stride, interleaved != 0);
  }
  return B200_OK;
}

// ===========================================================================
// KV slot write / gather (pure copies, bit exact)
// ===========================================================================
template <int ESZ, bool GATHER>
__global__ void __launch_bounds__(256) kv_copy_kernel(const int32_t* __restrict__ slot_ids,
                                                      const uint8_t* __restrict__ k_tok,
                                                      const uint8_t* __restrict__ v_tok,
                                                      uint8_t* __restrict__ k_cache,
                                                      uint8_t* __restrict__ v_cache,
                                                      int64_t row_bytes, int64_t k_stride_b,
                                                      int64_t v_stride_b, bool vec) {
  // GATHER: tok <- cache ; else cache <- tok.  (k_tok/v_tok are written when GATHER)
  const int64_t tok = blockIdx.x;
  const int64_t slot = slot_ids[tok];
  uint8_t* kc = k_cache + slot * row_bytes;
  uint8_t* vc = v_cache + slot * row_bytes;
  uint8_t* kt = const_cast<uint8_t*>(k_tok) + tok * k_stride_b;
  uint8_t* vt = const_cast<uint8_t*>(v_tok) + tok * v_stride_b;
  if (vec) {
    for (int64_t o = threadIdx.x * 16; o < row_bytes; o += blockDim.x * 16) {
      if (GATHER) {
        st_v4(kt + o, ld_v4(kc + o));
        st_v4(vt + o, ld_v4(vc + o));
      } else {
        st_v4(kc + o, ld_nc_v4(kt + o));
        st_v4(vc + o, ld_nc_v4(vt + o));
      }
    }
  } else {
    for (int64_t o = threadIdx.x * ESZ; o < row_bytes; o += blockDim.x * ESZ) {
#pragma unroll
      for (int b = 0; b < ESZ; ++b) {
        if (GATHER) {
          kt[o + b] = kc[o + b];
          vt[o + b] = vc[o + b];
        } else {
          kc[o + b] = kt[o + b];
          vc[o + b] = vt[o + b];
        }
      }
    }
  }
}

static int esize(int dtype) { return dtype == B200_FP32 ? 4 : 2; }

// ===========================================================================
// SiLU, SiLU*mul
// ===========================================================================
template <typename T>
__device__ __forceinline__ float silu_t(float x) {
  // (T)( x / (1 + __expf(-x)) )   (activation_kernels.cu:44-50)
  return rnd<T>(x / (1.0f + __expf(-x)));
}
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ACT 0 = SiLU, 1 = GELU "new" (tanh form), 2 = GELU "fast": fp32 math on the T-typed input, one
// rounding to T at the end; the expressions are shaped like the reference's so that the compiler
// contracts the same multiply-adds (activation_kernels.cu:13-50)
template <typename T, int ACT>
__device__ __forceinline__ float act_t(float x) {
  if constexpr (ACT == 0) {
    return silu_t<T>(x);
  } else if constexpr (ACT == 1) {
    const float cdf = 0.5f * (1.0f + tanh_approx((0.7978845608028654f * (x + 0.044715f * x * x * x))));
    return rnd<T>(x * cdf);
  } else {
    const float cdf = 0.5f * (1.0f + tanh_approx((0.7978845608028654f * x) * (1.0f + 0.044715f * x * x)));
    return rnd<T>(x * cdf);
  }
}

// MODE 0: out = act(a)      MODE 1: out = act(a) * b  (second rounding)
template <typename T, int MODE, int ACT = 0>
__global__ void __launch_bounds__(256) silu_kernel(T* __restrict__ out, const T* __restrict__ a,
                                                   const T* __restrict__ b, int64_t rows,
                                                   int64_t n, int64_t a_stride, int64_t b_stride,
                                                   bool vec) {
  constexpr int VEC = 16 / sizeof(T);
  pdl_wait();
  pdl_launch_dependents();
  if (vec) {
    const int64_t nv = n / VEC;
    const int64_t total = rows * nv;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / nv, j = idx % nv;
      uint4 ar = ld_nc_v4(a + r * a_stride + j * VEC);
      const T* ae = reinterpret_cast<const T*>(&ar);
      uint4 br;
      if (MODE == 1) br = ld_nc_v4(b + r * b_stride + j * VEC);
      const T* be = reinterpret_cast<const T*>(&br);
      uint4 orr;
      T* o = reinterpret_cast<T*>(&orr);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float s = act_t<T, ACT>(Num<T>::to_f(ae[i]));
        if (MODE == 1) s = s * Num<T>::to_f(be[i]);
        o[i] = Num<T>::from_f(s);
      }
      st_v4(out + r * n + j * VEC, orr);
    }
  } else {
    const int64_t total = rows * n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / n, j = idx % n;
      float s = act_t<T, ACT>(Num<T>::to_f(a[r * a_stride + j]));
      if (MODE == 1) s = s * Num<T>::to_f(b[r * b_stride + j]);
      out[r * n + j] = Num<T>::from_f(s);
    }
  }
}

// out = silu(gate) * up with gate | up = the two halves of the gate_up GEMM's row, delivered as
// that GEMM's stream-K partials: each is first rounded to T (the GEMM epilogue's rounding), then
// exactly the MODE 1 arithmetic above.
template <typename T>
__global__ void __launch_bounds__(256, 3) silu_mul_splitk_kernel(T* __restrict__ out,
                                                              const float* __restrict__ partials,
                                                              W4Plan plan, int64_t slot_stride,
                                                              int64_t rows, int inter) {
  const int nv = inter / 8;
  const int64_t total = rows * nv;
  // the first item's index arithmetic and contributor counts do not depend on the GEMM: before the wait
  const int64_t idx0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int j0 = (int)(idx0 % nv);
  const int cg0 = w4_contrib_col(plan, j0 * 8), cu0 = w4_contrib_col(plan, inter + j0 * 8);
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t idx = idx0; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / nv;
    const int j = (int)(idx - r * nv);
    const float* prow = partials + r * (2 * (int64_t)inter);
    float g[8], u[8];
    const bool first = idx == idx0;
    w4_sum_partials8(g, prow + j * 8, slot_stride, first ? cg0 : w4_contrib_col(plan, j * 8));
    w4_sum_partials8(u, prow + inter + j * 8, slot_stride, first ? cu0 : w4_contrib_col(plan, inter + j * 8));
    uint4 orr;
    T* o = reinterpret_cast<T*>(&orr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = silu_t<T>(rnd<T>(g[i])) * rnd<T>(u[i]);
      o[i] = Num<T>::from_f(s);
    }
    st_v4(out + r * inter + j * 8, orr);
  }
}

template <typename T, int MODE, int ACT = 0>
static int launch_silu(void* out, const void* a, const void* b, int64_t rows, int64_t n,
                       int64_t a_stride, int64_t b_stride, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  if (rows * n == 0) return B200_OK;
  bool vec = n % VEC == 0 && a_stride % VEC == 0 && is_aligned(out, 16) && is_aligned(a, 16);
  if (MODE == 1) vec = vec && b_stride % VEC == 0 && is_aligned(b, 16);
  const int64_t work = vec ? rows * (n / VEC) : rows * n;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  B200_PDL_LAUNCH("activation", (silu_kernel<T, MODE, ACT>), (unsigned)blocks, 256, 0, st, static_cast<T*>(out),
                  static_cast<const T*>(a), static_cast<const T*>(b), rows, n, a_stride, b_stride,
                  vec);
  return B200_OK;
}

// ===========================================================================
// greedy sampling tail: argmax over the vocabulary
// ===========================================================================
// torch.argmax semantics: first index of the maximum, NaN counts as the maximum.
__device__ __forceinline__ bool argmax_better(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an != bn) return an;
  if (an) return ia < ib;
  return a > b || (a == b && ia < ib);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_cluster_u32(const void* local_smem, uint32_t rank) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem)), ra, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}

// CL == 1: one block per row.  CL > 1 (long rows — the vocabulary): a thread-block cluster of CL
// blocks per row, block c takes the c-th share of the row's 16-byte vectors (the last one also the
// scalar tail); block 0 of the cluster picks the winner out of the CL candidates through
// distributed shared memory.  One block per row left 84 of the 148 SMs idle and took 26.6 us for
// [64, 128256] bf16 (16 MB: 2.5 us at HBM rate).
template <typename T, int CL>
__global__ void __launch_bounds__(512) argmax_kernel(int64_t* __restrict__ out,
                                                     const T* __restrict__ x, int n,
                                                     int64_t stride) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float sv[16];
  __shared__ int si[16];
  __shared__ uint32_t cand[2];   // this block's candidate: value bits, index
  pdl_wait();
  pdl_launch_dependents();
  const int c = CL > 1 ? (int)cluster_ctarank() : 0;
  const T* row = x + (int64_t)(blockIdx.x / CL) * stride;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const bool vec = (stride % VEC == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int nv_all = vec ? n / VEC : 0;
  const int per = (nv_all + CL - 1) / CL;
  const int v_begin = c * per, nv = min(nv_all, v_begin + per);
  for (int v = v_begin + threadIdx.x; v < nv; v += blockDim.x) {
    const uint4 raw = ld_nc_v4(row + v * VEC);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float f = Num<T>::to_f(e[i]);
      if (argmax_better(f, v * VEC + i, best, bi)) { best = f; bi = v * VEC + i; }
    }
  }
  if (c == CL - 1) {
    for (int j = nv_all * VEC + threadIdx.x; j < n; j += blockDim.x) {
      const float f = Num<T>::to_f(row[j]);
      if (argmax_better(f, j, best, bi)) { best = f; bi = j; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (argmax_better(ob, oi, best, bi)) { best = ob; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (int)(blockDim.x >> 5) ? sv[lane] : -INFINITY;
    bi = lane < (int)(blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (argmax_better(ob, oi, best, bi)) { best = ob; bi = oi; }
    }
    if (CL == 1) {
      if (lane == 0) out[blockIdx.x] = bi;
    } else if (lane == 0) {
      cand[0] = __float_as_uint(best);
      cand[1] = (uint32_t)bi;
    }
  }
  if constexpr (CL > 1) {
    cluster_sync_all();               // every block's candidate is in its shared memory
    if (c == 0 && warp == 0) {
      best = -INFINITY;
      bi = 0x7fffffff;
      if (lane < CL) {
        best = __uint_as_float(ld_cluster_u32(&cand[0], (uint32_t)lane));
        bi = (int)ld_cluster_u32(&cand[1], (uint32_t)lane);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (argmax_better(ob, oi, best, bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) out[blockIdx.x / CL] = bi;
    }
    cluster_sync_all();               // nobody leaves before block 0 has read its candidate
  }
}


}  // namespace b200

// ===========================================================================
// C ABI
// ===========================================================================
using namespace b200;

#define DISPATCH_DTYPE3(dtype, ...)                                    \
  switch (dtype) {                                                     \
    case B200_BF16: { using T = __nv_bfloat16; return __VA_ARGS__; }   \
    case B200_FP16: { using T = __half; return __VA_ARGS__; }          \
    case B200_FP32: { using T = float; return __VA_ARGS__; }           \
    default: return set_error(B200_ERR_INVALID_ARG, "bad dtype %d", dtype); \
  }
#define DISPATCH_DTYPE2(dtype, ...)                                    \
  switch (dtype) {                                                     \
    case B200_BF16: { using T = __nv_bfloat16; return __VA_ARGS__; }   \
    case B200_FP16: { using T = __half; return __VA_ARGS__; }          \
    default: return set_error(B200_ERR_UNSUPPORTED, "dtype %d not supported here", dtype); \
  }

template <typename T>
static int launch_argmax(int64_t* out, const T* x, int64_t rows, int64_t n, int64_t stride, cudaStream_t st) {
  constexpr int CL = 8;
  if (n >= 16384 && rows * CL <= 65535) {   // long rows: a cluster of 8 blocks per row
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(rows * CL));
    cfg.blockDim = dim3(512);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_level() >= 2 ? 2 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, argmax_kernel<T, CL>, out, x, (int)n, stride);
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "launch of argmax failed: %s", cudaGetErrorString(e));
    count_launch();
    return B200_OK;
  }
  B200_PDL_LAUNCH("argmax", (argmax_kernel<T, 1>), (unsigned)rows, 512, 0, st, out, x, (int)n, stride);
  return B200_OK;
}

template <typename T>
static int launch_row_norm(int which, void* out, const void* in, const void* weight, const void* bias,
                           int64_t rows, int64_t n, float eps, cudaStream_t st) {
  const int threads = (int)(n < 1024 ? ((n + 31) / 32) * 32 : 1024);
  if (which == 0) {
    B200_PDL_LAUNCH("gemma_rms_norm", gemma_rms_norm_kernel<T>, (unsigned)rows, threads, 0, st,
                    static_cast<T*>(out), static_cast<const T*>(in), static_cast<const T*>(weight), eps, n);
  } else {
    B200_PDL_LAUNCH("layer_norm", layer_norm_kernel<T>, (unsigned)rows, threads, 0, st, static_cast<T*>(out),
                    static_cast<const T*>(in), static_cast<const T*>(weight), static_cast<const T*>(bias),
                    eps, n);
  }
  return B200_OK;
}

extern "C" {

int b200_rms_norm(void* out, const void* in, const void* weight, int64_t rows, int64_t n,
                  float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in && weight, "rms_norm: null pointer");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n < (1ll << 31), "rms_norm: bad shape [%lld,%lld]",
                 (long long)rows, (long long)n);
  if (rows == 0) return B200_OK;
  DISPATCH_DTYPE3(dtype, (launch_rms_norm<T, false>(out, nullptr, in, weight, rows, n, eps,
                                                    static_cast<cudaStream_t>(stream))));
}

int b200_gemma_rms_norm(void* out, const void* in, const void* weight, int64_t rows, int64_t n,
                        float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in && weight, "gemma_rms_norm: null pointer");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n < (1ll << 31), "gemma_rms_norm: bad shape");
  if (rows == 0) return B200_OK;
  DISPATCH_DTYPE3(dtype, (launch_row_norm<T>(0, out, in, weight, nullptr, rows, n, eps,
                                             static_cast<cudaStream_t>(stream))));
}

int b200_layer_norm(void* out, const void* in, const void* weight, const void* bias, int64_t rows,
                    int64_t n, float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in && weight, "layer_norm: null pointer (bias may be NULL)");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n < (1ll << 31), "layer_norm: bad shape");
  if (rows == 0) return B200_OK;
  DISPATCH_DTYPE3(dtype, (launch_row_norm<T>(1, out, in, weight, bias, rows, n, eps,
                                             static_cast<cudaStream_t>(stream))));
}

int b200_rms_norm_residual(void* out, void* residual, const void* in, const void* weight,
                           int64_t rows, int64_t n, float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in && weight && residual, "rms_norm_residual: null pointer");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n < (1ll << 31), "rms_norm_residual: bad shape");
  if (rows == 0) return B200_OK;
  DISPATCH_DTYPE3(dtype, (launch_rms_norm<T, true>(out, residual, in, weight, rows, n, eps,
                                                   static_cast<cudaStream_t>(stream))));
}

int b200_rms_norm_residual_splitk(void* out, void* residual, const float* partials, int splits,
                                  int64_t gemm_k, const void* weight, int64_t rows, int64_t n,
                                  float eps, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && residual && partials && weight, "rms_norm_residual_splitk: null pointer");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n % 128 == 0 && n <= 32768 && gemm_k > 0 && gemm_k % 128 == 0,
                 "rms_norm_residual_splitk: need n %% 128 == 0, n <= 32768, gemm_k %% 128 == 0");
  const W4Plan plan = w4_get_plan(n, gemm_k, rows);
  B200_CHECK_ARG(splits == plan.slots,
                 "rms_norm_residual_splitk: partials of a [K=%lld, N=%lld] GEMM have %d slots, got %d",
                 (long long)gemm_k, (long long)n, plan.slots, splits);
  B200_CHECK_ARG(is_aligned(out, 16) && is_aligned(residual, 16) && is_aligned(partials, 16) &&
                     is_aligned(weight, 16),
                 "rms_norm_residual_splitk: 16-byte alignment required");
  if (rows == 0) return B200_OK;
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t stride = rows * n;
  switch (dtype) {
    case B200_BF16:
      return launch_rms_norm_splitk<__nv_bfloat16>(out, residual, partials, plan, stride, weight,
                                                   rows, n, eps, st);
    case B200_FP16:
      return launch_rms_norm_splitk<__half>(out, residual, partials, plan, stride, weight, rows, n,
                                            eps, st);
    default:
      return set_error(B200_ERR_UNSUPPORTED, "rms_norm_residual_splitk: bf16 / fp16 only");
  }
}

int b200_rope_inplace(void* q, void* k, const int32_t* positions, const void* cos_sin,
                      int64_t n_tokens, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim,
                      int64_t rotary_dim, int64_t q_stride, int64_t k_stride, int interleaved,
                      int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(q && k && positions && cos_sin, "rope: null pointer");
  B200_CHECK_ARG(rotary_dim > 0 && rotary_dim % 2 == 0 && rotary_dim <= head_dim,
                 "rope: rotary_dim %lld invalid for head_dim %lld", (long long)rotary_dim,
                 (long long)head_dim);
  B200_CHECK_ARG(q_stride >= n_heads * head_dim && k_stride >= n_kv_heads * head_dim,
                 "rope: heads must be dense within a token");
  DISPATCH_DTYPE3(dtype, (launch_rope<T, false>(q, k, nullptr, positions, cos_sin, nullptr, nullptr,
                                                nullptr, n_tokens, n_heads, n_kv_heads, head_dim,
                                                rotary_dim, q_stride, k_stride, 0, interleaved,
                                                static_cast<cudaStream_t>(stream))));
}

int b200_rope_kv_write(void* q, void* k, const void* v, const int32_t* positions,
                       const void* cos_sin, const int32_t* slot_ids, void* k_cache, void* v_cache,
                       int64_t n_tokens, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim,
                       int64_t rotary_dim, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                       int interleaved, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(q && k && v && positions && cos_sin && slot_ids && k_cache && v_cache,
                 "rope_kv_write: null pointer");
  B200_CHECK_ARG(rotary_dim > 0 && rotary_dim % 2 == 0 && rotary_dim <= head_dim,
                 "rope_kv_write: bad rotary_dim");
  B200_CHECK_ARG(q_stride >= n_heads * head_dim && k_stride >= n_kv_heads * head_dim &&
                     v_stride >= n_kv_heads * head_dim,
                 "rope_kv_write: heads must be dense within a token");
  DISPATCH_DTYPE3(dtype, (launch_rope<T, true>(q, k, v, positions, cos_sin, slot_ids, k_cache,
                                               v_cache, n_tokens, n_heads, n_kv_heads, head_dim,
                                               rotary_dim, q_stride, k_stride, v_stride,
                                               interleaved, static_cast<cudaStream_t>(stream))));
}

int b200_rope_kv_write_splitk(void* qkv, const float* partials, int splits, int64_t gemm_k,
                              const int32_t* positions, const void* cos_sin,
                              const int32_t* slot_ids, void* k_cache, void* v_cache,
                              int64_t n_tokens, int64_t n_heads, int64_t n_kv_heads,
                              int64_t head_dim, int64_t rotary_dim, int interleaved, int dtype,
                              b200_stream_t stream) {
  B200_CHECK_ARG(qkv && partials && positions && cos_sin && slot_ids && k_cache && v_cache,
                 "rope_kv_write_splitk: null pointer");
  B200_CHECK_ARG(rotary_dim > 0 && rotary_dim % 2 == 0 && rotary_dim <= head_dim,
                 "rope_kv_write_splitk: bad rotary_dim");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "rope_kv_write_splitk: bf16 / fp16 only");
  const int64_t n = (n_heads + 2 * n_kv_heads) * head_dim;
  B200_CHECK_ARG(n % 128 == 0 && gemm_k > 0 && gemm_k % 128 == 0 && n_tokens >= 0 && n_tokens <= 128,
                 "rope_kv_write_splitk: row of %lld is not a W4A16 GEMM output", (long long)n);
  if (n_tokens == 0) return B200_OK;
  RopePartials parts{};
  parts.plan = w4_get_plan(n, gemm_k, n_tokens);
  B200_CHECK_ARG(splits == parts.plan.slots, "rope_kv_write_splitk: expected %d partial slots, got %d",
                 parts.plan.slots, splits);
  parts.data = partials;
  parts.slot_stride = n_tokens * n;
  parts.row_n = (int)n;
  const int es = 2;
  uint8_t* base = static_cast<uint8_t*>(qkv);
  void* q = base;
  void* k = base + n_heads * head_dim * es;
  void* v = base + (n_heads + n_kv_heads) * head_dim * es;
  if (dtype == B200_BF16)
    return launch_rope<__nv_bfloat16, true>(q, k, v, positions, cos_sin, slot_ids, k_cache, v_cache,
                                            n_tokens, n_heads, n_kv_heads, head_dim, rotary_dim, n, n,
                                            n, interleaved, static_cast<cudaStream_t>(stream), &parts);
  return launch_rope<__half, true>(q, k, v, positions, cos_sin, slot_ids, k_cache, v_cache, n_tokens,
                                   n_heads, n_kv_heads, head_dim, rotary_dim, n, n, n, interleaved,
                                   static_cast<cudaStream_t>(stream), &parts);
}

int b200_silu_mul_splitk(void* out, const float* partials, int splits, int64_t gemm_k, int64_t rows,
                         int64_t inter, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && partials, "silu_mul_splitk: null pointer");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "silu_mul_splitk: bf16 / fp16 only");
  B200_CHECK_ARG(rows >= 0 && rows <= 128 && inter > 0 && inter % 64 == 0 && gemm_k > 0 &&
                     gemm_k % 128 == 0 && is_aligned(out, 16) && is_aligned(partials, 16),
                 "silu_mul_splitk: bad shape / alignment");
  if (rows == 0) return B200_OK;
  const W4Plan plan = w4_get_plan(2 * inter, gemm_k, rows);
  B200_CHECK_ARG(splits == plan.slots, "silu_mul_splitk: expected %d partial slots, got %d",
                 plan.slots, splits);
  const int64_t work = rows * (inter / 8);
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 3;   // one wave at 3 resident blocks per SM (grid-stride loop)
  if (blocks > cap) blocks = cap;
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_BF16)
    B200_PDL_LAUNCH_L(1, "silu_mul_splitk", silu_mul_splitk_kernel<__nv_bfloat16>, (unsigned)blocks, 256, 0,
                    st, static_cast<__nv_bfloat16*>(out), partials, plan, rows * 2 * inter, rows,
                    (int)inter);
  else
    B200_PDL_LAUNCH_L(1, "silu_mul_splitk", silu_mul_splitk_kernel<__half>, (unsigned)blocks, 256, 0, st,
                    static_cast<__half*>(out), partials, plan, rows * 2 * inter, rows, (int)inter);
  return B200_OK;
}

static int kv_copy(bool gather, const int32_t* slot_ids, const void* k, const void* v,
                   void* k_cache, void* v_cache, int64_t n_tokens, int64_t n_kv_heads,
                   int64_t head_dim, int64_t k_stride, int64_t v_stride, int dtype,
                   b200_stream_t stream) {
  B200_CHECK_ARG(slot_ids && k && v && k_cache && v_cache, "kv copy: null pointer");
  B200_CHECK_ARG(dtype >= 0 && dtype <= 2, "kv copy: bad dtype");
  B200_CHECK_ARG(k_stride >= n_kv_heads * head_dim && v_stride >= n_kv_heads * head_dim,
                 "kv copy: heads must be dense within a token");
  if (n_tokens == 0) return B200_OK;
  const int es = esize(dtype);
  const int64_t row_bytes = n_kv_heads * head_dim * es;
  const bool vec = row_bytes % 16 == 0 && (k_stride * es) % 16 == 0 && (v_stride * es) % 16 == 0 &&
                   is_aligned(k, 16) && is_aligned(v, 16) && is_aligned(k_cache, 16) &&
                   is_aligned(v_cache, 16);
  auto st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)n_tokens);
  const uint8_t* kb = static_cast<const uint8_t*>(k);
  const uint8_t* vb = static_cast<const uint8_t*>(v);
  uint8_t* kc = static_cast<uint8_t*>(k_cache);
  uint8_t* vc = static_cast<uint8_t*>(v_cache);
  if (es == 2) {
    if (gather)
      kv_copy_kernel<2, true><<<grid, 128, 0, st>>>(slot_ids, kb, vb, kc, vc, row_bytes,
                                                    k_stride * es, v_stride * es, vec);
    else
      kv_copy_kernel<2, false><<<grid, 128, 0, st>>>(slot_ids, kb, vb, kc, vc, row_bytes,
                                                     k_stride * es, v_stride * es, vec);
  } else {
    if (gather)
      kv_copy_kernel<4, true><<<grid, 128, 0, st>>>(slot_ids, kb, vb, kc, vc, row_bytes,
                                                    k_stride * es, v_stride * es, vec);
    else
      kv_copy_kernel<4, false><<<grid, 128, 0, st>>>(slot_ids, kb, vb, kc, vc, row_bytes,
                                                     k_stride * es, v_stride * es, vec);
  }
  B200_LAUNCH_OK("kv_copy");
  return B200_OK;
}

int b200_kv_write(const int32_t* slot_ids, const void* k, const void* v, void* k_cache,
                  void* v_cache, int64_t n_tokens, int64_t n_kv_heads, int64_t head_dim,
                  int64_t k_stride, int64_t v_stride, int dtype, b200_stream_t stream) {
  return kv_copy(false, slot_ids, k, v, k_cache, v_cache, n_tokens, n_kv_heads, head_dim, k_stride,
                 v_stride, dtype, stream);
}

int b200_kv_gather(const int32_t* slot_ids, const void* k_cache, const void* v_cache, void* k_out,
                   void* v_out, int64_t n_tokens, int64_t n_kv_heads, int64_t head_dim, int dtype,
                   b200_stream_t stream) {
  return kv_copy(true, slot_ids, k_out, v_out, const_cast<void*>(k_cache),
                 const_cast<void*>(v_cache), n_tokens, n_kv_heads, head_dim,
                 n_kv_heads * head_dim, n_kv_heads * head_dim, dtype, stream);
}

int b200_silu(void* out, const void* in, int64_t rows, int64_t n, int64_t in_stride, int dtype,
              b200_stream_t stream) {
  B200_CHECK_ARG(out && in, "silu: null pointer");
  B200_CHECK_ARG(rows >= 0 && n >= 0 && in_stride >= n, "silu: bad shape");
  DISPATCH_DTYPE3(dtype, (launch_silu<T, 0>(out, in, nullptr, rows, n, in_stride, 0,
                                            static_cast<cudaStream_t>(stream))));
}

// act: 1 = gelu_new, 2 = gelu_fast (activation_kernels.cu:22-41,108-119,128-145); with_mul: the input is
// [rows, 2n] and out = act(in[:, :n]) * in[:, n:] (second rounding), else in is [rows, n] with row
// stride in_stride
int b200_gelu(void* out, const void* in, int64_t rows, int64_t n, int64_t in_stride, int act,
              int with_mul, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in, "gelu: null pointer");
  B200_CHECK_ARG(rows >= 0 && n >= 0 && (act == 1 || act == 2), "gelu: bad shape / act (1 = new, 2 = fast)");
  auto st = static_cast<cudaStream_t>(stream);
  if (with_mul) {
    const void* up = static_cast<const uint8_t*>(in) + n * esize(dtype);
    if (act == 1) { DISPATCH_DTYPE3(dtype, (launch_silu<T, 1, 1>(out, in, up, rows, n, 2 * n, 2 * n, st))); }
    DISPATCH_DTYPE3(dtype, (launch_silu<T, 1, 2>(out, in, up, rows, n, 2 * n, 2 * n, st)));
  }
  B200_CHECK_ARG(in_stride >= n, "gelu: bad stride");
  if (act == 1) { DISPATCH_DTYPE3(dtype, (launch_silu<T, 0, 1>(out, in, nullptr, rows, n, in_stride, 0, st))); }
  DISPATCH_DTYPE3(dtype, (launch_silu<T, 0, 2>(out, in, nullptr, rows, n, in_stride, 0, st)));
}

int b200_silu_mul(void* out, const void* in, int64_t rows, int64_t n, int dtype,
                  b200_stream_t stream) {
  B200_CHECK_ARG(out && in, "silu_mul: null pointer");
  B200_CHECK_ARG(rows >= 0 && n >= 0, "silu_mul: bad shape");
  const int es = esize(dtype);
  const void* up = static_cast<const uint8_t*>(in) + n * es;
  DISPATCH_DTYPE3(dtype, (launch_silu<T, 1>(out, in, up, rows, n, 2 * n, 2 * n,
                                            static_cast<cudaStream_t>(stream))));
}

int b200_silu_mul_strided(void* out, const void* gate, const void* up, int64_t rows, int64_t n,
                          int64_t gate_stride, int64_t up_stride, int dtype,
                          b200_stream_t stream) {
  B200_CHECK_ARG(out && gate && up, "silu_mul_strided: null pointer");
  B200_CHECK_ARG(rows >= 0 && n >= 0 && gate_stride >= n && up_stride >= n,
                 "silu_mul_strided: bad shape");
  DISPATCH_DTYPE3(dtype, (launch_silu<T, 1>(out, gate, up, rows, n, gate_stride, up_stride,
                                            static_cast<cudaStream_t>(stream))));
}

int b200_argmax(int64_t* out, const void* logits, int64_t rows, int64_t n, int64_t stride, int dtype,
                b200_stream_t stream) {
  if (rows == 0) return B200_OK;  // an empty selection (prefill chunks before the last) has no storage
  B200_CHECK_ARG(out && logits, "argmax: null pointer");
  B200_CHECK_ARG(rows >= 0 && n > 0 && n < (1ll << 31) && stride >= n, "argmax: bad shape");
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE3(dtype, (launch_argmax<T>(out, static_cast<const T*>(logits), rows, n, stride, st)));
}

}  // extern "C"
