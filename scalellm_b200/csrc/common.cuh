// common.cuh — shared host/device helpers for libb200decode (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/b200_decode.h"

namespace b200 {

// ---------------------------------------------------------------------------
// host: error reporting (thread-local message, negative status codes)
// ---------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define B200_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return ::b200::set_error(B200_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define B200_CUDA_OK(expr)                                                     \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess)                                                     \
      return ::b200::set_error(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,  \
                               cudaGetErrorString(_e), __FILE__, __LINE__);    \
  } while (0)

// Checks the launch itself (bad config / missing image); never synchronises.
#define B200_LAUNCH_OK(name)                                                    \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess)                                                      \
      return ::b200::set_error(B200_ERR_CUDA, "launch of %s failed: %s", name,  \
                               cudaGetErrorString(_e));                         \
    ::b200::count_launch();                                                     \
  } while (0)

inline bool is_aligned(const void* p, size_t a) {
  return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

int sm_count();  // cached multiProcessorCount of the current device
long long* debug_trace_ptr();  // b200_debug_set_trace buffer (nullptr = off): per-CTA timestamps

// Programmatic dependent launch.  B200_PDL = 0: off; 1 (default): the W4A16 GEMM and the kernels
// that consume its partials are launched with the programmatic-stream-serialization attribute:
// the GEMM's CTAs become resident and prefetch weights while the predecessor kernel (which calls
// pdl_launch_dependents() at its start) is still running, and the GEMM lets its consumer's CTAs
// be scheduled when its last epilogue begins; 2: every hot kernel is launched that way.  Each
// kernel starts with pdl_wait(), which returns once the predecessor has completed and flushed.
int pdl_level();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int min_level, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                              size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_level() >= min_level ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// launch + check + count; `level` = the B200_PDL level from which the launch is programmatic
#define B200_PDL_LAUNCH(name, kernel, grid, block, smem, st, ...) \
  B200_PDL_LAUNCH_L(2, name, kernel, grid, block, smem, st, __VA_ARGS__)
#define B200_PDL_LAUNCH_L(level, name, kernel, grid, block, smem, st, ...)                     \
  do {                                                                                         \
    cudaError_t _e =                                                                           \
        ::b200::launch_pdl(level, kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__);     \
    if (_e != cudaSuccess)                                                                     \
      return ::b200::set_error(B200_ERR_CUDA, "launch of %s failed: %s", name,                 \
                               cudaGetErrorString(_e));                                        \
    ::b200::count_launch();                                                                    \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no libcuda link).
typedef CUresult (*tensor_map_encode_fn)(
    CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
    CUtensorMapFloatOOBfill);
tensor_map_encode_fn get_tensor_map_encode();

// ---------------------------------------------------------------------------
// device: numeric helpers
// ---------------------------------------------------------------------------
template <typename T>
struct Num;

template <>
struct Num<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
  // packed pair (lo, hi) -> two floats
  static __device__ __forceinline__ float2 unpack(uint32_t p) {
    float2 r;
    r.x = __uint_as_float(p << 16);
    r.y = __uint_as_float(p & 0xffff0000u);
    return r;
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

template <>
struct Num<__half> {
  using T2 = __half2;
  static __device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
  static __device__ __forceinline__ __half from_f(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float2 unpack(uint32_t p) {
    return __half22float2(*reinterpret_cast<__half2*>(&p));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

template <>
struct Num<float> {
  static __device__ __forceinline__ float to_f(float x) { return x; }
  static __device__ __forceinline__ float from_f(float x) { return x; }
};

// round-trip through T: the reference computes "in T" by converting to float,
// doing ONE op, and rounding back (c10::BFloat16 / c10::Half operators).
template <typename T>
__device__ __forceinline__ float rnd(float x) {
  return Num<T>::to_f(Num<T>::from_f(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the THREADS threads of the CTA (red: 32 floats of shared memory); every thread gets it
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  constexpr int NW = THREADS / 32;
  float t = (lane < NW) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}

// Sum of squares of a row whose n fp32 values are staged in shared memory `sq`, in EXACTLY the
// order of the reference's rms_norm kernels, so that the normalised row is bit-identical to theirs
// for every dtype (src/kernels/layernorm_kernels.cu:30-35,137-145 + reduce_kernel_utils.cuh:15-64):
// their block has BD = min(n, 1024) threads; thread t accumulates x[t], x[t+BD], ... with one FFMA
// each (variance += x * x, contracted), every warp of 32 consecutive t does an xor-butterfly
// (16, 8, 4, 2, 1), lane 0 parks the warp's sum, and one more butterfly over the <= 32 warp sums
// (zeros for absent warps) gives the total.  Our CTA has THREADS threads whatever n is: real warp
// w plays the reference's warps w, w + THREADS/32, ...  Every thread returns the total.
// `red`: 32 floats of shared memory.  Contains the __syncthreads() that publish `sq`.
template <int THREADS>
__device__ __forceinline__ float row_sumsq_ref_order(const float* sq, float* red, int n) {
  __syncthreads();
  const int BD = n < 1024 ? n : 1024;
  const int nvw = (BD + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int vw = warp; vw < nvw; vw += THREADS / 32) {
    const int t = vw * 32 + lane;
    float v = 0.f;
    if (t < BD)
      for (int i = t; i < n; i += BD) v = fmaf(sq[i], sq[i], v);
    v = warp_sum(v);
    if (lane == 0) red[vw] = v;
  }
  __syncthreads();
  return warp_sum(lane < nvw ? red[lane] : 0.f);
}

// 128-bit streaming global access
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_v4(const void* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void st_v4(void* p, uint4 v) {
  *reinterpret_cast<uint4*>(p) = v;
}

// ---------------------------------------------------------------------------
// device: mbarrier / TMA / tcgen05 PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 3-D tiled TMA load, completion on an mbarrier of this CTA
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both 16-B aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 --------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A operand read from tensor memory (lane = row,
// 8 x 32-bit columns = 16 bf16 of K per MMA)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 registers -> 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when done
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile whose rows are 128 B (64 bf16) wide:
// 8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major,
// descriptor version 1 (Blackwell), layout_type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored)  [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version = 1    [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B   [61,64)
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, MxN tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)            // c_format = F32
         | (1u << 7)          // a_format = BF16
         | (1u << 10)         // b_format = BF16
         | (0u << 15)         // a_major  = K
         | (0u << 16)         // b_major  = K
         | ((N >> 3) << 17)   // n_dim
         | ((M >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------------------
// W4A16 stream-K partition, shared by the GEMM and by every kernel that consumes its fp32
// partials.  The (n tile, k tile) units of C = A * W (k fastest) are cut into P equal contiguous
// shares, CTA p owns [w4_unit_begin(p), w4_unit_begin(p+1)).  A tile nt is covered by the
// consecutive CTAs first..last; CTA p writes its part of tile nt to partial slot (p - first), so
// a consumer sums slots [0, w4_contrib(nt)) of that tile's columns — no atomics, no fix-up pass.
// An "n tile" of the partition is 128 << nsub_log2 output columns (two weight tiles share one
// activation stage when the batch fits 64 rows, halving the activation traffic out of L2).
// ---------------------------------------------------------------------------
struct W4Plan {
  int units, P, KT, NT, slots;  // slots = max contributors of any tile (partials buffer depth)
  int nsub_log2;                // 128-column weight tiles per n tile of the partition: 1 << nsub_log2
};
__host__ __device__ __forceinline__ int w4_unit_begin(int p, int units, int P) {
  return (int)(((long long)p * units) / P);
}
// CTA that owns unit u: the largest p with begin(p) <= u.  floor(p * units / P) <= u  <=>
// p * units < (u + 1) * P  <=>  p <= floor(((u + 1) * P - 1) / units) — exact, one division (the
// consumers evaluate this per 32-byte group of a row: 32-bit arithmetic, units * P < 2^31 always:
// units <= 2^16 tiles per GEMM would need a 1M x 1M weight).
__host__ __device__ __forceinline__ int w4_owner(int u, int units, int P) {
  const unsigned p = ((unsigned)(u + 1) * (unsigned)P - 1u) / (unsigned)units;
  return p > (unsigned)(P - 1) ? P - 1 : (int)p;
}
__host__ __device__ __forceinline__ int w4_first_owner(const W4Plan& pl, int nt) {
  return w4_owner(nt * pl.KT, pl.units, pl.P);
}
__host__ __device__ __forceinline__ int w4_contrib(const W4Plan& pl, int nt) {
  return w4_owner(nt * pl.KT + pl.KT - 1, pl.units, pl.P) - w4_owner(nt * pl.KT, pl.units, pl.P) + 1;
}
// contributors of the partition tile that holds output column `col`
__host__ __device__ __forceinline__ int w4_contrib_col(const W4Plan& pl, int col) {
  return w4_contrib(pl, col >> (7 + pl.nsub_log2));
}
constexpr int W4_MAX_SLOTS = 8;
// The partition b200_w4a16_gemm_splitk uses for M rows x a [K, N] weight on the current device.
W4Plan w4_get_plan(int64_t N, int64_t K, int64_t M);
// partials [slots][M][N] -> C (bf16, + bias): the reduction pass of the plain GEMM entry points
// (w4a16.cu; N % 8 == 0, the partition's tiles may overhang N)
int w4_launch_reduce(__nv_bfloat16* C, const float* partials, const __nv_bfloat16* bias, int M, int N,
                     int64_t ldc, int64_t slot_stride, const W4Plan& plan, cudaStream_t st);

// prefill_attn.cu: the tcgen05 prefill / chunked-prefill attention kernel behind
// b200_paged_attn_decode (taken for >= 64 packed query rows at head_dim 128)
bool prefill_attn_eligible(int max_q_len, int group, int head_dim, int block_size);
int launch_prefill_attn(void* out, const void* q, const void* k_cache, const void* v_cache,
                        const int32_t* q_cu_lens, const int32_t* kv_cu_lens, const int32_t* block_table,
                        const int32_t* block_cu_lens, const float* alibi, int64_t batch, int64_t n_tokens_bound,
                        int n_heads, int n_kv_heads, int64_t n_slots, int64_t q_stride_t, int64_t q_stride_h,
                        int64_t o_stride_t, int64_t o_stride_h, int64_t kv_stride_s, int64_t kv_stride_h,
                        int block_size, int max_q_len, float sm_scale, float soft_cap, int window, int dtype,
                        cudaStream_t st);

// Device side of programmatic dependent launch: pdl_wait() blocks until the predecessor kernel
// has completed and its writes are visible; pdl_launch_dependents() allows the successor's early
// start (both are no-ops for an ordinary launch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Sum of the partial slots of 8 consecutive columns (one 32-byte group inside one n tile) of
// row `row_ptr`: the loads of the `count` contributing slots are all issued up front (one L2 round
// trip; absent slots are neither loaded — they may hold stale NaNs — nor added), fixed summation
// order ((p0 + p1) + p2) + ...  Round 1 loaded all 8 slots with the absent ones clamped onto the
// last one: 4x the L2 traffic at the usual 2 contributors — the gate_up consumer spent 14.7 us
// on it (profiles/r02_step_timeline.md).
__device__ __forceinline__ void w4_sum_partials8(float (&a)[8], const float* __restrict__ p0,
                                                 int64_t slot_stride, int count) {
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  // four slots per round: their loads are in flight together (one L2 round trip for the usual
  // 1-4 contributors) at 32 registers instead of 64 — the SiLU*mul consumer, which sums two such
  // groups, ran at 178 registers = one 256-thread block per SM, three waves (r02 ncu capture)
#pragma unroll
  for (int s0 = 0; s0 < W4_MAX_SLOTS; s0 += 4) {
    if (s0 < count) {
      float4 lo[4], hi[4];
#pragma unroll
      for (int sp = 0; sp < 4; ++sp) {
        if (s0 + sp < count) {
          const float4* src = reinterpret_cast<const float4*>(p0 + (s0 + sp) * slot_stride);
          lo[sp] = __ldcg(src);
          hi[sp] = __ldcg(src + 1);
        }
      }
#pragma unroll
      for (int sp = 0; sp < 4; ++sp) {
        if (s0 + sp < count) {
          a[0] += lo[sp].x; a[1] += lo[sp].y; a[2] += lo[sp].z; a[3] += lo[sp].w;
          a[4] += hi[sp].x; a[5] += hi[sp].y; a[6] += hi[sp].z; a[7] += hi[sp].w;
        }
      }
    }
  }
}

}  // namespace b200
