// w4a16.cu — int4-weight x bf16-activation matmul for sm_100a.
//
// Replaces marlin::{awq,gptq}_repack + marlin::gptq_gemm
// (src/kernels/quantization/marlin.h:17-37) behind the qlinear plugins.
//
//   C[M,N] = A[M,K] * W,  W[k,n] = bf16_mul(bf16(q) - bf16(z), s)     (exact sub,
//   one rounding in the multiply — marlin/numeric_conversion.h:144-167,221-240),
//   fp32 accumulation in TMEM, fp32 cross-CTA reduction, one final rounding.
//
// B200 design (not Marlin's mma.sync/cp.async pipeline):
//   * operands swapped so the WEIGHT tile is the 128-row UMMA "A" operand and the
//     decode batch (M <= 128) is the UMMA N dimension: D[n, m] in TMEM.
//   * weights live in HBM as self-contained "tile blobs" (128 n x 128 k: packed
//     nibbles + that tile's scales + zero points), streamed with one bulk-async
//     copy per blob into a deep shared-memory ring;
//   * four groups of 4 dequant warps take k-tiles round-robin: a thread owns one weight row (one
//     TMEM lane), turns its 16 packed words into 128 bf16 (lop3 magic-number int4->bf16,
//     HSUB2 zero point, HMUL2 scale) and writes them with tcgen05.st straight into TENSOR
//     MEMORY, where they are the A operand of the UMMA — no shared-memory staging, no
//     swizzle arithmetic and no generic->async proxy fence on the weight path;
//   * one thread issues tcgen05.mma (kind::f16, A from TMEM, M=128, N=MT, K=16) with the
//     accumulator in TMEM (double buffered); activations arrive by TMA (128B swizzle);
//   * stream-K over (n_tile, k_tile) units so all SMs stream an equal share of
//     the weight bytes; every CTA writes the fp32 partial of each of its segments to a slot of
//     the partials buffer (slot = its rank among the tile's contributors, common.cuh W4Plan)
//     and the CONSUMER of the GEMM sums the slots in a fixed order (deterministic): the
//     following RMSNorm, the TP all-reduce's copy-in, or w4_reduce_kernel for a plain bf16 C.
//     No atomics, no fix-up tail in the GEMM.
//
// W4 tile blob (n_tile nt, k_tile kt) at ((nt * K/128) + kt) * blob_bytes:
//   [0, 8192)            qdata: uint4[(khalf*2+q4)*128 + n_local]; word w of that
//                        uint4 holds k = kt*128 + (khalf*8+q4*4+w)*8 + {0..7} of
//                        column nt*128+n_local, nibble p<4 -> k+2p, p>=4 -> k+2(p-4)+1
//   [8192, +ngrp*256)    scales bf16 [ngrp][128]     (ngrp = 128/geff, geff = min(g,128))
//   [.., +ngrp*128)      zero points, uint8 [ngrp][128] (0..16: GPTQ-v1 "zero+1" can reach 16)

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

constexpr int W4_BN = 128;
constexpr int W4_BK = 128;
constexpr int W4_QBYTES = 8192;
constexpr int W4_MAX_BLOB = 8192 + 4 * (256 + 128);  // 9728

__host__ __device__ inline int w4_geff(int g) { return (g <= 0 || g > 128) ? 128 : g; }
__host__ __device__ inline int w4_blob_bytes(int geff) {
  return W4_QBYTES + (128 / geff) * (256 + 128);
}

// ===========================================================================
// prepack: checkpoint format -> tile blobs
// ===========================================================================
// MODE 0 = AWQ (qweight [K, N/8], nibble order [0,2,4,6,1,3,5,7] along N)
// MODE 1 = GPTQ (qweight [K/8, N], natural nibble order along K)
template <int MODE>
__device__ __forceinline__ uint32_t load_q(const int32_t* __restrict__ qweight, int64_t k,
                                           int64_t n, int64_t N) {
  if (MODE == 0) {
    const uint32_t w = (uint32_t)qweight[k * (N / 8) + (n >> 3)];
    const int c = (int)(n & 7);
    const int pos = (c >> 1) + ((c & 1) << 2);  // inverse of [0,2,4,6,1,3,5,7]
    return (w >> (4 * pos)) & 0xf;
  } else {
    const uint32_t w = (uint32_t)qweight[(k >> 3) * N + n];
    return (w >> (4 * (int)(k & 7))) & 0xf;
  }
}

template <int MODE>
__device__ __forceinline__ uint32_t load_z(const int32_t* __restrict__ qzeros, int64_t grp,
                                           int64_t n, int64_t N, int plus_one) {
  if (qzeros == nullptr) return 8u;  // symmetric GPTQ (qlinear_gptq_marlin_impl.cpp:18-20)
  const uint32_t w = (uint32_t)qzeros[grp * (N / 8) + (n >> 3)];
  const int c = (int)(n & 7);
  const int pos = MODE == 0 ? (c >> 1) + ((c & 1) << 2) : c;
  return ((w >> (4 * pos)) & 0xf) + (plus_one ? 1u : 0u);  // may be 16 (qlinear_impl.cpp:44)
}

// The 8192 nibble bytes of tile (nt, kt): 256 threads, thread <-> (row n_local, k half).  `perm`
// (GPTQ act-order, else nullptr): row k of the packed weight is row perm[k] of the checkpoint —
// the rows sorted by group (marlin/gptq_repack.cu:17 does the same gather on its way into the
// Marlin layout; qlinear_gptq_marlin_impl.cpp:43-56 builds perm = argsort(g_idx)).
template <int MODE>
__device__ __forceinline__ void w4_pack_tile_nibbles(uint8_t* __restrict__ out_tile,
                                                     const int32_t* __restrict__ qweight,
                                                     const int32_t* __restrict__ perm, int kt,
                                                     int nt, int64_t N) {
  const int t = threadIdx.x;
  const int n_local = t & 127, khalf = t >> 7;
  const int64_t n = (int64_t)nt * 128 + n_local;
#pragma unroll
  for (int q4 = 0; q4 < 2; ++q4) {
    uint32_t words[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int64_t k0 = (int64_t)kt * 128 + (khalf * 8 + q4 * 4 + w) * 8;
      uint32_t word = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int dk = p < 4 ? 2 * p : 2 * (p - 4) + 1;
        const int64_t k = perm ? (int64_t)perm[k0 + dk] : k0 + dk;
        word |= load_q<MODE>(qweight, k, n, N) << (4 * p);
      }
      words[w] = word;
    }
    uint4 v = make_uint4(words[0], words[1], words[2], words[3]);
    *reinterpret_cast<uint4*>(out_tile + ((khalf * 2 + q4) * 128 + n_local) * 16) = v;
  }
}

// g_sorted (act-order, else nullptr): quant group of packed row k = g_sorted[k] (= g_idx[perm[k]],
// non-decreasing; every aligned block of min(g, 128) packed rows lies inside one group because all
// groups have exactly g rows when the whole K is present — is_k_full).
template <int MODE>
__global__ void __launch_bounds__(256) w4_prepack_kernel(uint8_t* __restrict__ packed,
                                                         const int32_t* __restrict__ qweight,
                                                         const int32_t* __restrict__ qzeros,
                                                         const __nv_bfloat16* __restrict__ scales,
                                                         const int32_t* __restrict__ perm,
                                                         const int32_t* __restrict__ g_sorted,
                                                         int64_t K, int64_t N, int g_actual,
                                                         int geff, int plus_one) {
  const int kt = blockIdx.x, nt = blockIdx.y;
  const int KT = (int)(K / W4_BK);
  const int ngrp = 128 / geff;
  const int blob = w4_blob_bytes(geff);
  uint8_t* out = packed + ((int64_t)nt * KT + kt) * blob;
  const int t = threadIdx.x;
  w4_pack_tile_nibbles<MODE>(out, qweight, perm, kt, nt, N);
  __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(out + W4_QBYTES);
  uint8_t* z_out = out + W4_QBYTES + ngrp * 256;
  for (int i = t; i < ngrp * 128; i += 256) {
    const int grp = i >> 7, nl = i & 127;
    const int64_t k_first = (int64_t)kt * 128 + grp * geff;
    const int64_t gi = g_sorted ? (int64_t)g_sorted[k_first] : k_first / g_actual;
    s_out[i] = scales[gi * N + (int64_t)nt * 128 + nl];
    z_out[i] = (uint8_t)load_z<MODE>(qzeros, gi, (int64_t)nt * 128 + nl, N, plus_one);
  }
}

// ---------------------------------------------------------------------------------------------
// The operator-level drop-in (shim/b200_kernels: marlin::awq_repack / gptq_repack / gptq_gemm with
// the reference's EXACT signatures, src/kernels/quantization/marlin.h:17-37).  There the repack
// sees only q_weight and must fill an `out` of the same byte count ((K/16) x (N*16/8) int32 =
// K*N/2 bytes), and scales / zero points reach gptq_gemm separately, already in Marlin's column
// order (the layer permuted them: qlinear_awq_marlin_impl.cpp:62-124).  So:
//   repack:   `out` = the nibble part of every tile blob, tile (nt, kt) at byte (nt*KT + kt) * 8192;
//   assemble: nibble tiles + Marlin-order scales (+ Marlin-packed zero points, or the symmetric 8)
//             -> the full tile blobs the GEMM streams (done once per weight, cached by the shim).
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) w4_repack_nibbles_kernel(uint8_t* __restrict__ out,
                                                                const int32_t* __restrict__ qweight,
                                                                const int32_t* __restrict__ perm,
                                                                int64_t K, int64_t N) {
  const int kt = blockIdx.x, nt = blockIdx.y;
  const int KT = (int)(K / W4_BK);
  w4_pack_tile_nibbles<MODE>(out + ((int64_t)nt * KT + kt) * W4_QBYTES, qweight, perm, kt, nt, N);
}

// Marlin's column order of scales / zero points (tests/kernels/quant_utils.py:231-241,282-291):
// inside every block of 64 columns natural column c sits at position (c % 8) * 8 + c / 8; with a
// single group, inside every block of 32 columns at ((c % 8) / 2) * 8 + 2 * (c / 8) + c % 2.
__device__ __forceinline__ int64_t marlin_scale_pos(int64_t n, bool single_group) {
  if (!single_group) {
    const int c = (int)(n & 63);
    return (n & ~63ll) + (c & 7) * 8 + (c >> 3);
  }
  const int c = (int)(n & 31);
  return (n & ~31ll) + ((c & 7) >> 1) * 8 + 2 * (c >> 3) + (c & 1);
}
// zero points: the 64-column permutation above, then nibble t of a packed word holds position
// [0,2,4,6,1,3,5,7][t] of its group of 8 (qlinear_awq_marlin_impl.cpp:84-96)
__device__ __forceinline__ uint32_t marlin_zero_at(const int32_t* __restrict__ zeros, int64_t gi,
                                                   int64_t n, int64_t N) {
  const int64_t pos = marlin_scale_pos(n, false);
  const uint32_t w = (uint32_t)zeros[gi * (N / 8) + (pos >> 3)];
  const int u = (int)(pos & 7);
  const int t = (u >> 1) + ((u & 1) << 2);  // inverse of [0,2,4,6,1,3,5,7]
  return (w >> (4 * t)) & 0xf;
}

__global__ void __launch_bounds__(256) w4_assemble_marlin_kernel(
    uint8_t* __restrict__ packed, const uint8_t* __restrict__ nibbles,
    const __nv_bfloat16* __restrict__ scales_m, const int32_t* __restrict__ zeros_m, int64_t K,
    int64_t N, int g_actual, int geff) {
  const int kt = blockIdx.x, nt = blockIdx.y;
  const int KT = (int)(K / W4_BK);
  const int ngrp = 128 / geff;
  uint8_t* out = packed + ((int64_t)nt * KT + kt) * w4_blob_bytes(geff);
  const uint4* src = reinterpret_cast<const uint4*>(nibbles + ((int64_t)nt * KT + kt) * W4_QBYTES);
  const int t = threadIdx.x;
  reinterpret_cast<uint4*>(out)[t] = src[t];
  reinterpret_cast<uint4*>(out)[t + 256] = src[t + 256];
  __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(out + W4_QBYTES);
  uint8_t* z_out = out + W4_QBYTES + ngrp * 256;
  const bool single = g_actual >= K;
  for (int i = t; i < ngrp * 128; i += 256) {
    const int grp = i >> 7, nl = i & 127;
    const int64_t gi = ((int64_t)kt * 128 + grp * geff) / g_actual;
    const int64_t n = (int64_t)nt * 128 + nl;
    s_out[i] = scales_m[gi * N + marlin_scale_pos(n, single)];
    z_out[i] = zeros_m ? (uint8_t)marlin_zero_at(zeros_m, gi, n, N) : (uint8_t)8;
  }
}

// out[r, j] = in[r, perm[j]] (2-byte elements): the activation side of GPTQ act-order
// (permute_cols_kernel, marlin/gptq_gemm.cu:66-104)
__global__ void __launch_bounds__(256) permute_cols_kernel(uint16_t* __restrict__ out,
                                                           const uint16_t* __restrict__ in,
                                                           const int32_t* __restrict__ perm,
                                                           int cols, int64_t in_stride,
                                                           int64_t out_stride) {
  pdl_wait();
  pdl_launch_dependents();
  const uint16_t* src = in + (int64_t)blockIdx.x * in_stride;
  uint16_t* dst = out + (int64_t)blockIdx.x * out_stride;
  for (int j = threadIdx.x; j < cols; j += 256) dst[j] = src[perm[j]];
}

// ===========================================================================
// the dequant arithmetic shared by every consumer of a blob
// ===========================================================================
// one packed word (8 nibbles, pair-interleaved) -> 8 consecutive-k bf16 weights
__device__ __forceinline__ uint4 w4_dequant_word(uint32_t q, __nv_bfloat162 zmagic,
                                                 __nv_bfloat162 s2) {
  uint32_t r[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint32_t v;
    // (q & 0x000f000f) | 0x43004300  ->  bf16x2 (128 + nibble_j, 128 + nibble_{j+4})
    asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(v) : "r"(q >> (4 * jj)));
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    b = __hsub2(b, zmagic);  // exact: (128+q) - (128+z)
    b = __hmul2(b, s2);      // the single bf16 rounding
    r[jj] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ __nv_bfloat162 w4_zmagic(uint32_t z) {
  const uint32_t m = 0x4300u | z;  // bf16(128 + z)
  const uint32_t mm = m | (m << 16);
  return *reinterpret_cast<const __nv_bfloat162*>(&mm);
}

// blob -> dense W[K,N] bf16 (debug / bit-exact prepack parity)
__global__ void __launch_bounds__(256) w4_dequant_kernel(__nv_bfloat16* __restrict__ w_out,
                                                         const uint8_t* __restrict__ packed,
                                                         int64_t K, int64_t N, int geff) {
  const int kt = blockIdx.x, nt = blockIdx.y;
  const int KT = (int)(K / W4_BK);
  const int ngrp = 128 / geff;
  const uint8_t* blob = packed + ((int64_t)nt * KT + kt) * w4_blob_bytes(geff);
  const __nv_bfloat16* s_in = reinterpret_cast<const __nv_bfloat16*>(blob + W4_QBYTES);
  const uint8_t* z_in = blob + W4_QBYTES + ngrp * 256;
  const int t = threadIdx.x, n_local = t & 127, khalf = t >> 7;
#pragma unroll
  for (int q4 = 0; q4 < 2; ++q4) {
    const uint4 u = *reinterpret_cast<const uint4*>(blob + ((khalf * 2 + q4) * 128 + n_local) * 16);
    const uint32_t words[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int kc = khalf * 8 + q4 * 4 + w;
      const int grp = (kc * 8) / geff;
      const __nv_bfloat16 s = s_in[grp * 128 + n_local];
      const uint32_t z = z_in[grp * 128 + n_local];
      const uint4 d = w4_dequant_word(words[w], w4_zmagic(z), __halves2bfloat162(s, s));
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&d);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w_out[((int64_t)kt * 128 + kc * 8 + i) * N + (int64_t)nt * 128 + n_local] = e[i];
    }
  }
}

// ===========================================================================
// bring-up / debug GEMM on CUDA cores (selected only by B200_W4A16_IMPL=simt)
// ===========================================================================
// grid (N/128, ceil(M/8)); thread = one output column, 8 rows.
__global__ void __launch_bounds__(128) w4_gemm_simt_kernel(
    __nv_bfloat16* __restrict__ C, const __nv_bfloat16* __restrict__ A,
    const uint8_t* __restrict__ packed, const __nv_bfloat16* __restrict__ bias, int M, int N,
    int K, int64_t lda, int64_t ldc, int geff) {
  const int nt = blockIdx.x, m0 = blockIdx.y * 8, n_local = threadIdx.x;
  const int KT = K / W4_BK, ngrp = 128 / geff, blob_bytes = w4_blob_bytes(geff);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int kt = 0; kt < KT; ++kt) {
    const uint8_t* blob = packed + ((int64_t)nt * KT + kt) * blob_bytes;
    const __nv_bfloat16* s_in = reinterpret_cast<const __nv_bfloat16*>(blob + W4_QBYTES);
    const uint8_t* z_in = blob + W4_QBYTES + ngrp * 256;
    for (int kc = 0; kc < 16; ++kc) {
      const int khalf = kc >> 3, q4 = (kc >> 2) & 1, w = kc & 3;
      const uint32_t word =
          reinterpret_cast<const uint32_t*>(blob + ((khalf * 2 + q4) * 128 + n_local) * 16)[w];
      const int grp = (kc * 8) / geff;
      const __nv_bfloat16 s = s_in[grp * 128 + n_local];
      const uint32_t z = z_in[grp * 128 + n_local];
      const uint4 d = w4_dequant_word(word, w4_zmagic(z), __halves2bfloat162(s, s));
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&d);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < M) {
          const __nv_bfloat16* a = A + (int64_t)(m0 + r) * lda + kt * 128 + kc * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            acc[r] = fmaf(__bfloat162float(a[i]), __bfloat162float(e[i]), acc[r]);
        }
      }
    }
  }
  const int n = nt * 128 + n_local;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (m0 + r < M) {
      __nv_bfloat16 o = __float2bfloat16_rn(acc[r]);
      if (bias) o = __float2bfloat16_rn(__bfloat162float(o) + __bfloat162float(bias[n]));
      C[(int64_t)(m0 + r) * ldc + n] = o;
    }
  }
}

// ===========================================================================
// tcgen05 stream-K GEMM (partials out)
// ===========================================================================
template <int MT, int NSUB>
struct W4Cfg {
  static constexpr int ACT_STAGES = MT <= 64 ? 6 : 3;   // activation ring (L2 / TMA latency)
  static constexpr int RAW_STAGES = MT <= 64 ? 11 : 10; // weight-blob ring (HBM latency)
  // accumulators [128 x MT] fp32: double buffered (epilogue of a segment overlaps the next
  // segment's MMAs) for one weight tile per unit; single buffered for two, so that the
  // dequantised-weight ring keeps 6 slots (the MMA warp then waits for the epilogue to drain at
  // a segment boundary, which a CTA crosses at most a couple of times)
  static constexpr int ACC_BUFS = (NSUB == 1 && MT <= 64) || MT > 64 ? 2 : 1;
  static constexpr int ACC_COLS = ACC_BUFS * NSUB * MT;
  static constexpr int A_COL0 = ACC_COLS < 128 ? 128 : ACC_COLS;
  static constexpr int A_STAGES = (512 - A_COL0) / 64;  // dequantised-weight slots in TMEM: 6 or 4
  static constexpr int ACT_ATOM = MT * 128;       // bytes of one [MT x 64] bf16 swizzle atom
  static constexpr int ACT_BYTES = 2 * ACT_ATOM;  // 128 k per stage
  static constexpr int RAW_BYTES = W4_MAX_BLOB;   // 9728 = 76 * 128
  static constexpr int TMEM_COLS = 512;
  static constexpr int N_BARS = 2 * RAW_STAGES + 2 * ACT_STAGES + 2 * A_STAGES + 4;
  static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)ACT_STAGES * ACT_BYTES +
                                 (size_t)RAW_STAGES * RAW_BYTES + N_BARS * 8 + 64;
  static_assert(ACC_COLS <= A_COL0 && A_COL0 + A_STAGES * 64 <= TMEM_COLS && A_STAGES >= 4,
                "TMEM over-subscribed");
};

// warp roles: 0-15 dequant (4 groups x 4 lane quadrants), 16 weight-blob producer, 17 activation
// producer, 18 MMA issuer, 19 idle, 20-23 epilogue (lane quadrant = warp % 4)
constexpr int W4_DEQ_GROUPS = 4;
constexpr int W4_DEQ_WARPS = 4 * W4_DEQ_GROUPS;
constexpr int W4_WARP_RAW = 16, W4_WARP_ACT = 17, W4_WARP_MMA = 18, W4_WARP_EPI = 20;
constexpr int W4_THREADS = 24 * 32;

struct W4Params {
  const uint8_t* packed;
  float* partials;        // [slots][M][N] fp32
  int64_t slot_stride;    // M * N
  int M, N, KT, geff_log2, ngrp, blob_bytes;
  W4Plan plan;
  long long* trace;       // debug: [grid][16] clock64 milestones (null in production)
};

#define W4_TRACE(slot)                                                               \
  do {                                                                               \
    if constexpr (TRACE) p.trace[(int64_t)blockIdx.x * 16 + (slot)] = clock64();     \
  } while (0)
// slots 14 / 15: %globaltimer (ns, common to all SMs) at CTA entry / exit: the launch ramp and the
// span of the whole grid, which the per-CTA clock64 milestones cannot show
#define W4_TRACE_NS(slot)                                                            \
  do {                                                                               \
    if constexpr (TRACE) {                                                           \
      unsigned long long ns_;                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_));                        \
      p.trace[(int64_t)blockIdx.x * 16 + (slot)] = (long long)ns_;                   \
    }                                                                                \
  } while (0)

struct SegIter {
  int u, u1, KT;
  __device__ __forceinline__ bool next(int& nt, int& kt0, int& kt1) {
    if (u >= u1) return false;
    nt = u / KT;
    kt0 = u - nt * KT;
    kt1 = min(KT, kt0 + (u1 - u));
    u += kt1 - kt0;
    return true;
  }
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t r;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}

// NSUB weight tiles (adjacent n tiles, same k tile) share one activation stage: the activation
// bytes pulled out of L2 per weight tile drop by NSUB (the kernel is L2-bandwidth bound on them:
// every CTA re-reads the [MT x 128] activation tile of each of its k tiles).
//
// Round 2 measured ten variants of the MMA loop (one commit per tile, tiles issued in pairs, three
// dequant groups, one "full" barrier per stage, two issuing warps and their combinations): all
// within 3 % of this form on every projection shape (profiles/r02_w4_variants.md), so they are gone.
template <int MT, int NSUB, bool TRACE>
__global__ void __launch_bounds__(W4_THREADS, 1)
w4a16_gemm_kernel(const __grid_constant__ CUtensorMap amap, const W4Params p) {
  using Cfg = W4Cfg<MT, NSUB>;
  // A dequant group waits for its weight blob by the parity of the ring entry's use count.  That
  // is sound only while no group can get a whole ring ahead of a blob that has not landed; the
  // TMEM slot ring bounds a group's lead over the in-order MMA issuer to (groups + slots) tiles.
  static_assert(Cfg::RAW_STAGES > W4_DEQ_GROUPS + Cfg::A_STAGES, "weight ring too shallow for the slot ring");
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* act_smem = base;
  uint8_t* raw_smem = act_smem + Cfg::ACT_STAGES * Cfg::ACT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_smem + Cfg::RAW_STAGES * Cfg::RAW_BYTES);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + Cfg::RAW_STAGES;
  uint64_t* act_full = raw_empty + Cfg::RAW_STAGES;
  uint64_t* act_empty = act_full + Cfg::ACT_STAGES;
  uint64_t* deq_full = act_empty + Cfg::ACT_STAGES;
  uint64_t* deq_empty = deq_full + Cfg::A_STAGES;
  uint64_t* tmem_full = deq_empty + Cfg::A_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KT = p.KT;
  if (threadIdx.x == 0) {
    W4_TRACE(0);
    W4_TRACE_NS(14);
  }
  const int u_begin = w4_unit_begin(blockIdx.x, p.plan.units, p.plan.P);
  const int u_end = w4_unit_begin(blockIdx.x + 1, p.plan.units, p.plan.P);

  // The weight ring belongs to the producer warp: it initialises the ring's barriers itself and
  // puts the first RAW_STAGES blobs in flight BEFORE the CTA-wide set-up (TMEM allocation, the
  // other barriers, the __syncthreads below) — the first HBM request leaves at ~200 instead of
  // ~1500 cycles after the CTA starts (profiles/r02_w4_trace.md: setup_done 1470).  Weights are
  // never written by another kernel, so no griddepcontrol.wait either.  The dequant warps learn
  // of the initialised barriers through the __syncthreads.
  int raw_pre = 0;  // weight tiles already requested by the prologue (warp W4_WARP_RAW, lane 0)
  if (warp == W4_WARP_RAW && lane == 0) {
    for (int i = 0; i < Cfg::RAW_STAGES; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], 4);
    }
    fence_mbar_init();
    SegIter it{u_begin, u_end, KT};
    int nt, kt0, kt1;
    while (raw_pre < Cfg::RAW_STAGES && it.next(nt, kt0, kt1)) {
      for (int kt = kt0; kt < kt1 && raw_pre < Cfg::RAW_STAGES; ++kt)
        for (int sub = 0; sub < NSUB && raw_pre < Cfg::RAW_STAGES; ++sub, ++raw_pre) {
          mbar_arrive_expect_tx(&raw_full[raw_pre], (uint32_t)p.blob_bytes);
          bulk_load_1d(raw_smem + raw_pre * Cfg::RAW_BYTES,
                       p.packed + ((int64_t)(nt * NSUB + sub) * KT + kt) * p.blob_bytes,
                       (uint32_t)p.blob_bytes, &raw_full[raw_pre]);
        }
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::ACT_STAGES; ++i) {
      mbar_init(&act_full[i], 1);
      mbar_init(&act_empty[i], 1);
    }
    for (int i = 0; i < Cfg::A_STAGES; ++i) {
      mbar_init(&deq_full[i], 4);
      mbar_init(&deq_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == W4_WARP_MMA) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp == W4_WARP_ACT && lane == 0) prefetch_tensormap(&amap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) W4_TRACE(1);

  if (warp < W4_DEQ_WARPS) {
    // ===================== dequant warps =====================================
    // group = warp / 4 takes weight tiles cnt % 4 == group; tile cnt goes to TMEM slot
    // cnt % A_STAGES (a ring shared by the groups: with more slots than groups a group runs ahead
    // of the tensor pipe instead of waiting for its previous tile's MMAs).  Inside a group warp
    // q = warp % 4 owns TMEM lanes [32q, 32q+32): thread <-> weight row n_local, all 128 k.
    const int group = warp >> 2;
    const int n_local = (warp & 3) * 32 + lane;
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + Cfg::A_COL0;
    const uint32_t raw_u32 = smem_u32(raw_smem);
    const uint32_t sz_off = W4_QBYTES + n_local * 2;               // this row's scale, group 0
    const uint32_t zp_off = W4_QBYTES + p.ngrp * 256 + n_local;    // this row's zero point
    SegIter it{u_begin, u_end, KT};
    int nt, kt0, kt1, cnt = 0;  // cnt counts WEIGHT tiles (NSUB per unit)
    long long w_raw = 0, w_slot = 0;  // TRACE: cycles spent waiting
    while (it.next(nt, kt0, kt1)) {
      for (int kt = kt0; kt < kt1; ++kt)
      for (int sub = 0; sub < NSUB; ++sub, ++cnt) {
        if ((cnt & (W4_DEQ_GROUPS - 1)) != group) continue;
        const int rs = cnt % Cfg::RAW_STAGES;
        const int as = cnt % Cfg::A_STAGES;
        const uint32_t rph = (cnt / Cfg::RAW_STAGES) & 1, aph = (cnt / Cfg::A_STAGES) & 1;
        const uint32_t raw = raw_u32 + rs * Cfg::RAW_BYTES;
        const uint32_t a_tmem = a_lane + as * 64;
        long long tw = TRACE ? clock64() : 0;
        mbar_wait(&raw_full[rs], rph);
        if (TRACE) w_raw += clock64() - tw;
        if (TRACE && threadIdx.x == 0 && cnt == 0) W4_TRACE(2);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // 64 k = 32 TMEM columns per tcgen05.st
          uint4 u[2];
          __nv_bfloat162 s2[2], zm[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            // uint4 (hh, q) covers k in [32 (2hh+q), +32): quant group (32 (2hh+q)) >> log2(geff)
            const int qq = hh * 2 + q;
            u[q] = lds128(raw + (qq * 128 + n_local) * 16);
            const int gq = (qq * 32) >> p.geff_log2;
            const uint32_t sv = lds_u16(raw + sz_off + gq * 256);
            const uint32_t sv2 = sv | (sv << 16);
            s2[q] = *reinterpret_cast<const __nv_bfloat162*>(&sv2);
            zm[q] = w4_zmagic(lds_u8(raw + zp_off + gq * 128));
          }
          uint32_t r[32];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t words[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint4 d = w4_dequant_word(words[w], zm[q], s2[q]);
              r[(q * 4 + w) * 4 + 0] = d.x;
              r[(q * 4 + w) * 4 + 1] = d.y;
              r[(q * 4 + w) * 4 + 2] = d.z;
              r[(q * 4 + w) * 4 + 3] = d.w;
            }
          }
          if (hh == 0) {  // the MMAs that read this slot's previous tile must have drained
            tw = TRACE ? clock64() : 0;
            mbar_wait(&deq_empty[as], aph ^ 1);
            if (TRACE) w_slot += clock64() - tw;
            tc_fence_after();
          }
          tmem_st_32x32b_x32(a_tmem + hh * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&deq_full[as]);
          mbar_arrive(&raw_empty[rs]);
        }
      }
    }
    if (threadIdx.x == 0) W4_TRACE(3);
    if constexpr (TRACE) {
      if (threadIdx.x == 0) {
        p.trace[(int64_t)blockIdx.x * 16 + 12] = w_raw;
        p.trace[(int64_t)blockIdx.x * 16 + 13] = w_slot;
      }
    }
  } else if (warp == W4_WARP_RAW) {
    // ===================== weight-blob producer ==============================
    // Weights are never written by another kernel: under programmatic dependent launch this
    // warp starts streaming them while the predecessor kernel is still running (no pdl_wait).
    if (lane == 0) {
      SegIter it{u_begin, u_end, KT};
      int nt, kt0, kt1, cnt = 0;
      while (it.next(nt, kt0, kt1)) {
        for (int kt = kt0; kt < kt1; ++kt)
        for (int sub = 0; sub < NSUB; ++sub, ++cnt) {
          if (cnt < raw_pre) continue;  // requested by the prologue above
          const int rs = cnt % Cfg::RAW_STAGES;
          const uint32_t rph = (cnt / Cfg::RAW_STAGES) & 1;
          mbar_wait(&raw_empty[rs], rph ^ 1);
          mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)p.blob_bytes);
          bulk_load_1d(raw_smem + rs * Cfg::RAW_BYTES,
                       p.packed + ((int64_t)(nt * NSUB + sub) * KT + kt) * p.blob_bytes,
                       (uint32_t)p.blob_bytes, &raw_full[rs]);
        }
      }
    }
    __syncwarp();  // reconverge before the CTA-wide barrier at the end
  } else if (warp == W4_WARP_ACT) {
    // ===================== activation producer (TMA) =========================
    if (lane == 0) {
      pdl_wait();  // activations are the predecessor kernel's output
      SegIter it{u_begin, u_end, KT};
      int nt, kt0, kt1, cnt = 0;
      while (it.next(nt, kt0, kt1)) {
        for (int kt = kt0; kt < kt1; ++kt, ++cnt) {
          const int as = cnt % Cfg::ACT_STAGES;
          const uint32_t aph = (cnt / Cfg::ACT_STAGES) & 1;
          mbar_wait(&act_empty[as], aph ^ 1);
          uint64_t* full = &act_full[as];
          mbar_arrive_expect_tx(full, (uint32_t)Cfg::ACT_BYTES);
          uint8_t* dst = act_smem + as * Cfg::ACT_BYTES;
          tma_load_2d(dst, &amap, full, kt * 128, 0);
          tma_load_2d(dst + Cfg::ACT_ATOM, &amap, full, kt * 128 + 64, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == W4_WARP_MMA) {
    // ===================== MMA issuer =========================================
    // The whole warp runs this loop converged so every operand is warp-uniform; one elected lane
    // issues the UMMAs.  This one thread paces the kernel: it shares its scheduler with four
    // dequant warps, and every tcgen05.commit / barrier wait costs it ~170 cycles
    // (tools/microbench/mix.cu), so a weight tile takes ~600 cycles here against 256 of MMA time.
    // Tried and measured no better: releasing slots in pairs (one commit per two
    // tiles), two issuing warps with separate accumulators, loop bodies specialised on the ring
    // position.
    constexpr uint32_t idesc = umma_idesc_bf16(128, MT);
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t act_base = __shfl_sync(0xffffffffu, smem_u32(act_smem), 0);
    long long w_act = 0, w_deq = 0, w_acc = 0;  // TRACE: cycles spent waiting per barrier kind
    {
      SegIter it{u_begin, u_end, KT};
      int nt, kt0, kt1, cnt = 0, ucnt = 0, seg = 0;
      while (it.next(nt, kt0, kt1)) {
        const int buf = Cfg::ACC_BUFS == 2 ? (seg & 1) : 0;
        const uint32_t tph = (Cfg::ACC_BUFS == 2 ? (seg >> 1) : seg) & 1;
        long long tw = TRACE ? clock64() : 0;
        mbar_wait(&tmem_empty[buf], tph ^ 1);
        if (TRACE) w_acc += clock64() - tw;
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt, ++ucnt) {
          const int as = ucnt % Cfg::ACT_STAGES;
          const uint32_t aph = (ucnt / Cfg::ACT_STAGES) & 1;
          tw = TRACE ? clock64() : 0;
          mbar_wait(&act_full[as], aph);
          if (TRACE) w_act += clock64() - tw;
          const uint64_t b_desc0 = umma_desc_kmajor_sw128(act_base + as * Cfg::ACT_BYTES);
          const uint32_t first = (kt > kt0) ? 1u : 0u;
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub, ++cnt) {
            const int ds = cnt % Cfg::A_STAGES;
            const uint32_t dph = (cnt / Cfg::A_STAGES) & 1;
            const uint32_t a_tmem = tbase + Cfg::A_COL0 + ds * 64;
            const uint32_t d_tmem = tbase + (buf * NSUB + sub) * MT;
            tw = TRACE ? clock64() : 0;
            mbar_wait(&deq_full[ds], dph);
            if (TRACE) w_deq += clock64() - tw;
            if (TRACE && cnt == 0 && lane == 0) W4_TRACE(4);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                // descriptor start-address field is in 16-byte units: advance by constants
                const uint64_t b_desc =
                    b_desc0 + (uint64_t)(((ks >> 2) * Cfg::ACT_ATOM + (ks & 3) * 32) >> 4);
                umma_bf16_ts(d_tmem, a_tmem + ks * 8, b_desc, idesc, ks > 0 ? 1u : first);
              }
              umma_commit(&deq_empty[ds]);
              if (sub == NSUB - 1) {
                umma_commit(&act_empty[as]);
                if (kt == kt1 - 1) umma_commit(&tmem_full[buf]);
              }
            }
            __syncwarp();
          }
        }
        ++seg;
      }
    }
    if (lane == 0) W4_TRACE(5);
    if constexpr (TRACE) {
      if (lane == 0) {
        p.trace[(int64_t)blockIdx.x * 16 + 9] = w_act;
        p.trace[(int64_t)blockIdx.x * 16 + 10] = w_deq;
        p.trace[(int64_t)blockIdx.x * 16 + 11] = w_acc;
      }
    }
    __syncwarp();
  } else if (warp >= W4_WARP_EPI) {
    // ===================== epilogue warps (4) =================================
    // fp32 partial of every segment -> partials[slot][m][n]; slot = position of this CTA among
    // the tile's contributors.  Lane <-> n, so one store instruction writes 128 contiguous bytes.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int n_local = quad * 32 + lane;
    pdl_wait();  // the partials buffer may still be read by an earlier kernel's consumer
    SegIter it{u_begin, u_end, KT};
    int nt, kt0, kt1, seg = 0;
    while (it.next(nt, kt0, kt1)) {
      const int buf = Cfg::ACC_BUFS == 2 ? (seg & 1) : 0;
      const uint32_t tph = (Cfg::ACC_BUFS == 2 ? (seg >> 1) : seg) & 1;
      const int slot = (int)blockIdx.x - w4_first_owner(p.plan, nt);
      float* part = p.partials + (int64_t)slot * p.slot_stride + (int64_t)nt * (128 * NSUB) + n_local;
      mbar_wait(&tmem_full[buf], tph);
      // last segment of this CTA: let the consumer kernel's CTAs be scheduled now, so they are
      // resident (parked in pdl_wait) when this grid drains
      if (it.u >= it.u1 && warp == W4_WARP_EPI && lane == 0) pdl_launch_dependents();
      if (TRACE && warp == W4_WARP_EPI && lane == 0 && seg == 0) W4_TRACE(6);
      tc_fence_after();
      constexpr int CH = MT >= 32 ? 32 : 16;
#pragma unroll 1
      for (int sub = 0; sub < NSUB; ++sub) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * NSUB + sub) * MT;
#pragma unroll 1
        for (int c0 = 0; c0 < MT; c0 += CH) {
          uint32_t r[CH];
          if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, r);
          else tmem_ld_32x32b_x16(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int m = c0 + i;
            if (m < p.M) part[(int64_t)m * p.N + sub * 128] = __uint_as_float(r[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      ++seg;
    }
    if (warp == W4_WARP_EPI && lane == 0) W4_TRACE(7);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    W4_TRACE(8);
    W4_TRACE_NS(15);
  }
  if (warp == W4_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// partials -> C (bf16): the reduction pass of the plain GEMM entry point (callers that can,
// fuse this sum into their own first read instead: norm, all-reduce).
__global__ void __launch_bounds__(256) w4_reduce_kernel(__nv_bfloat16* __restrict__ C,
                                                        const float* __restrict__ partials,
                                                        const __nv_bfloat16* __restrict__ bias,
                                                        int M, int N, int64_t ldc,
                                                        int64_t slot_stride, W4Plan plan) {
  pdl_wait();
  pdl_launch_dependents();
  const int nvec = N / 8;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)M * nvec) return;
  const int m = (int)(idx / nvec), v = (int)(idx - (int64_t)m * nvec);
  const int count = w4_contrib_col(plan, v * 8);
  float a[8];
  w4_sum_partials8(a, partials + (int64_t)m * N + v * 8, slot_stride, count);
  __nv_bfloat16 o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] = __float2bfloat16_rn(a[i]);
    if (bias) o[i] = __float2bfloat16_rn(__bfloat162float(o[i]) + __bfloat162float(bias[v * 8 + i]));
  }
  __nv_bfloat16* dst = C + (int64_t)m * ldc + v * 8;
  if ((ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = o[i];
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct AMapKey {
  const void* ptr;
  int64_t M, K, lda;
  int mt;
  bool operator==(const AMapKey& o) const {
    return ptr == o.ptr && M == o.M && K == o.K && lda == o.lda && mt == o.mt;
  }
};
struct AMapEntry {
  AMapKey key;
  CUtensorMap map;
};
static std::mutex g_amap_mu;
static std::vector<AMapEntry> g_amaps;

static int get_act_tensor_map(const AMapKey& key, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_amap_mu);
    for (const auto& e : g_amaps)
      if (e.key == key) {
        *out = e.map;
        return B200_OK;
      }
  }
  tensor_map_encode_fn enc = get_tensor_map_encode();
  if (!enc) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)key.K, (cuuint64_t)key.M};
  cuuint64_t strides[1] = {(cuuint64_t)key.lda * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)key.mt};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(key.ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for activations", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_amap_mu);
    if (g_amaps.size() > 4096) g_amaps.clear();
    g_amaps.push_back({key, m});
  }
  *out = m;
  return B200_OK;
}

static long long* g_w4_trace = nullptr;
long long* debug_trace_ptr() { return g_w4_trace; }

static int pick_mt(int64_t M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128; }

template <int MT, int NSUB, bool TRACE>
static int launch_w4_kernel(const CUtensorMap& amap, const W4Params& p, cudaStream_t st) {
  using Cfg = W4Cfg<MT, NSUB>;
  auto kern = w4a16_gemm_kernel<MT, NSUB, TRACE>;
  B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM));
  B200_PDL_LAUNCH_L(1, "w4a16_gemm", kern, (unsigned)p.plan.P, W4_THREADS, Cfg::SMEM, st, amap, p);
  return B200_OK;
}

template <int MT>
static int launch_w4_gemm(const CUtensorMap& amap, const W4Params& p, cudaStream_t st) {
  if constexpr (MT <= 64) {
    if (p.plan.nsub_log2 == 1)
      return p.trace ? launch_w4_kernel<MT, 2, true>(amap, p, st)
                     : launch_w4_kernel<MT, 2, false>(amap, p, st);
  }
  return p.trace ? launch_w4_kernel<MT, 1, true>(amap, p, st)
                 : launch_w4_kernel<MT, 1, false>(amap, p, st);
}

// Stream-K partition of a [K, N] weight on this device: as many CTAs as SMs, but never so many
// that a tile has more than W4_MAX_SLOTS contributors (share >= KT / (W4_MAX_SLOTS - 1) units).
struct PlanKey {
  int64_t N, K;
  int nsub_log2, P;
};
static std::mutex g_plan_mu;
static std::vector<std::pair<PlanKey, W4Plan>> g_plans;

// pure arithmetic (no device, no environment): the partition for `ctas` CTAs
static W4Plan w4_make_plan(int64_t N, int64_t K, int nsub_log2, int ctas) {
  W4Plan pl{};
  pl.nsub_log2 = nsub_log2;
  pl.KT = (int)(K / 128);
  pl.NT = (int)(N / 128) >> nsub_log2;
  pl.units = pl.KT * pl.NT;
  const long long cap = (long long)(W4_MAX_SLOTS - 1) * pl.NT;
  int Pc = ctas;
  if (Pc > pl.units) Pc = pl.units;
  if (Pc > cap) Pc = (int)cap;
  if (Pc < 1) Pc = 1;
  while ((long long)pl.units * Pc >= (1ll << 31)) Pc /= 2;  // w4_owner's 32-bit arithmetic (never in practice)
  pl.P = Pc;
  pl.slots = 1;
  for (int nt = 0; nt < pl.NT; ++nt) {
    const int c = w4_contrib(pl, nt);
    if (c > pl.slots) pl.slots = c;
  }
  return pl;
}

W4Plan w4_get_plan(int64_t N, int64_t K, int64_t M) {
  int P = sm_count();
  const char* env = getenv("B200_W4A16_CTAS");
  if (env && atoi(env) > 0) P = atoi(env);
  if (P < 1) P = 1;
  const char* ens = getenv("B200_W4_NSUB");
  const int NT128 = (int)(N / 128);
  // B200_W4_NSUB=2: two weight tiles per activation stage (halves the activation traffic out of
  // L2).  Measured no faster than one tile per stage (the MMA-issuing thread, not L2, paces the
  // kernel), so it is opt-in.
  const bool want2 = ens && ens[0] == '2';
  const int nsub_log2 = (M <= 64 && NT128 % 2 == 0 && want2) ? 1 : 0;
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (const auto& e : g_plans)
      if (e.first.N == N && e.first.K == K && e.first.nsub_log2 == nsub_log2 && e.first.P == P)
        return e.second;
  }
  const W4Plan pl = w4_make_plan(N, K, nsub_log2, P);
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() > 256) g_plans.clear();
    g_plans.push_back({PlanKey{N, K, nsub_log2, P}, pl});
  }
  return pl;
}

int w4_launch_reduce(__nv_bfloat16* C, const float* partials, const __nv_bfloat16* bias, int M, int N,
                     int64_t ldc, int64_t slot_stride, const W4Plan& plan, cudaStream_t st) {
  const int64_t nvec = (int64_t)M * (N / 8);
  B200_PDL_LAUNCH_L(1, "w4a16_reduce", w4_reduce_kernel, (unsigned)((nvec + 255) / 256), 256, 0, st, C,
                    partials, bias, M, N, ldc, slot_stride, plan);
  return B200_OK;
}

static int log2_int(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// partial GEMM for one chunk of <= 128 rows
static int run_partial_gemm(float* partials, const void* A, const void* packed, int64_t M,
                            int64_t N, int64_t K, int64_t lda, int group_size, const W4Plan& plan,
                            cudaStream_t st) {
  const int geff = w4_geff(group_size);
  const int mt = pick_mt(M);
  W4Params p{};
  p.packed = static_cast<const uint8_t*>(packed);
  p.partials = partials;
  p.slot_stride = M * N;
  p.M = (int)M;
  p.N = (int)N;
  p.KT = plan.KT;
  p.geff_log2 = log2_int(geff);
  p.ngrp = 128 / geff;
  p.blob_bytes = w4_blob_bytes(geff);
  p.plan = plan;
  p.trace = g_w4_trace;
  CUtensorMap amap;
  AMapKey key{A, M, K, lda, mt};
  int rc = get_act_tensor_map(key, &amap);
  if (rc != B200_OK) return rc;
  switch (mt) {
    case 16: return launch_w4_gemm<16>(amap, p, st);
    case 32: return launch_w4_gemm<32>(amap, p, st);
    case 64: return launch_w4_gemm<64>(amap, p, st);
    default: return launch_w4_gemm<128>(amap, p, st);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_w4a16_packed_bytes(int64_t K, int64_t N, int group_size) {
  if (K <= 0 || N <= 0 || K % 128 || N % 128) return -1;
  return (K / 128) * (N / 128) * (int64_t)w4_blob_bytes(w4_geff(group_size));
}

static int check_w4_shape(const char* who, int64_t K, int64_t N, int g) {
  B200_CHECK_ARG(K > 0 && N > 0 && K % 128 == 0 && N % 128 == 0,
                 "%s: K=%lld and N=%lld must be positive multiples of 128", who, (long long)K,
                 (long long)N);
  B200_CHECK_ARG(g == -1 || g == 32 || g == 64 || g == 128 || (g > 128 && g % 128 == 0 && K % g == 0),
                 "%s: group_size %d not in {-1,32,64,128,k*128}", who, g);
  return B200_OK;
}

static int prepack(int mode, void* packed, const int32_t* qweight, const int32_t* qzeros,
                   const void* scales, int64_t K, int64_t N, int g, int plus_one,
                   b200_stream_t stream, const int32_t* perm = nullptr,
                   const int32_t* g_sorted = nullptr) {
  B200_CHECK_ARG(packed && qweight && scales, "w4a16_prepack: null pointer");
  B200_CHECK_ARG(mode == 1 || qzeros != nullptr, "w4a16_prepack_awq: qzeros required");
  int rc = check_w4_shape("w4a16_prepack", K, N, g);
  if (rc != B200_OK) return rc;
  B200_CHECK_ARG(is_aligned(packed, 16), "w4a16_prepack: packed buffer must be 16-byte aligned");
  const int g_actual = g <= 0 ? (int)K : g;
  const int geff = w4_geff(g);
  dim3 grid((unsigned)(K / 128), (unsigned)(N / 128));
  auto st = static_cast<cudaStream_t>(stream);
  if (mode == 0)
    w4_prepack_kernel<0><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(packed), qweight, qzeros,
                                               static_cast<const __nv_bfloat16*>(scales), nullptr,
                                               nullptr, K, N, g_actual, geff, 0);
  else
    w4_prepack_kernel<1><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(packed), qweight, qzeros,
                                               static_cast<const __nv_bfloat16*>(scales), perm,
                                               g_sorted, K, N, g_actual, geff, plus_one);
  B200_LAUNCH_OK("w4a16_prepack");
  return B200_OK;
}

int b200_w4a16_prepack_awq(void* packed, const int32_t* qweight, const int32_t* qzeros,
                           const void* scales, int64_t K, int64_t N, int group_size,
                           b200_stream_t stream) {
  return prepack(0, packed, qweight, qzeros, scales, K, N, group_size, 0, stream);
}

int b200_w4a16_prepack_gptq(void* packed, const int32_t* qweight, const int32_t* qzeros,
                            const void* scales, int64_t K, int64_t N, int group_size,
                            int zeros_plus_one, b200_stream_t stream) {
  return prepack(1, packed, qweight, qzeros, scales, K, N, group_size, zeros_plus_one, stream);
}

int b200_w4a16_prepack_gptq_actorder(void* packed, const int32_t* qweight, const int32_t* qzeros,
                                     const void* scales, const int32_t* perm,
                                     const int32_t* g_idx_sorted, int64_t K, int64_t N,
                                     int group_size, int zeros_plus_one, b200_stream_t stream) {
  B200_CHECK_ARG(perm && g_idx_sorted, "w4a16_prepack_gptq_actorder: perm and sorted g_idx required");
  B200_CHECK_ARG(group_size > 0, "w4a16_prepack_gptq_actorder: act-order needs a positive group size");
  return prepack(1, packed, qweight, qzeros, scales, K, N, group_size, zeros_plus_one, stream, perm,
                 g_idx_sorted);
}

static int repack_nibbles(int mode, void* out, const int32_t* qweight, const int32_t* perm, int64_t K,
                          int64_t N, b200_stream_t stream) {
  B200_CHECK_ARG(out && qweight, "w4a16_repack: null pointer");
  int rc = check_w4_shape("w4a16_repack", K, N, 128);
  if (rc != B200_OK) return rc;
  B200_CHECK_ARG(is_aligned(out, 16), "w4a16_repack: out must be 16-byte aligned");
  dim3 grid((unsigned)(K / 128), (unsigned)(N / 128));
  auto st = static_cast<cudaStream_t>(stream);
  if (mode == 0)
    w4_repack_nibbles_kernel<0><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(out), qweight, nullptr, K, N);
  else
    w4_repack_nibbles_kernel<1><<<grid, 256, 0, st>>>(static_cast<uint8_t*>(out), qweight, perm, K, N);
  B200_LAUNCH_OK("w4a16_repack");
  return B200_OK;
}

int b200_w4a16_repack_awq(void* out, const int32_t* qweight, int64_t K, int64_t N,
                          b200_stream_t stream) {
  return repack_nibbles(0, out, qweight, nullptr, K, N, stream);
}

int b200_w4a16_repack_gptq(void* out, const int32_t* qweight, const int32_t* perm, int64_t K,
                           int64_t N, b200_stream_t stream) {
  return repack_nibbles(1, out, qweight, perm, K, N, stream);
}

int b200_w4a16_assemble_marlin(void* packed, const void* nibbles, const void* scales_marlin,
                               const int32_t* zeros_marlin, int64_t K, int64_t N, int group_size,
                               b200_stream_t stream) {
  B200_CHECK_ARG(packed && nibbles && scales_marlin, "w4a16_assemble_marlin: null pointer");
  int rc = check_w4_shape("w4a16_assemble_marlin", K, N, group_size);
  if (rc != B200_OK) return rc;
  B200_CHECK_ARG(is_aligned(packed, 16) && is_aligned(nibbles, 16),
                 "w4a16_assemble_marlin: buffers must be 16-byte aligned");
  B200_CHECK_ARG(N % 64 == 0, "w4a16_assemble_marlin: N %% 64 == 0 (Marlin's column blocks)");
  const int g_actual = group_size <= 0 ? (int)K : group_size;
  dim3 grid((unsigned)(K / 128), (unsigned)(N / 128));
  w4_assemble_marlin_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint8_t*>(packed), static_cast<const uint8_t*>(nibbles),
      static_cast<const __nv_bfloat16*>(scales_marlin), zeros_marlin, K, N, g_actual,
      w4_geff(group_size));
  B200_LAUNCH_OK("w4a16_assemble_marlin");
  return B200_OK;
}

int b200_permute_cols(void* out, const void* in, const int32_t* perm, int64_t rows, int64_t cols,
                      int64_t in_stride, int64_t out_stride, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && in && perm, "permute_cols: null pointer");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "permute_cols: bf16 / fp16 only");
  B200_CHECK_ARG(rows >= 0 && cols > 0 && cols < (1ll << 31) && in_stride >= cols && out_stride >= cols,
                 "permute_cols: bad shape");
  B200_CHECK_ARG(out != in, "permute_cols: out must not alias in");
  if (rows == 0) return B200_OK;
  B200_PDL_LAUNCH("permute_cols", permute_cols_kernel, (unsigned)rows, 256, 0,
                  static_cast<cudaStream_t>(stream), static_cast<uint16_t*>(out),
                  static_cast<const uint16_t*>(in), perm, (int)cols, in_stride, out_stride);
  return B200_OK;
}

int b200_w4a16_dequant(void* w_out, const void* packed, int64_t K, int64_t N, int group_size,
                       b200_stream_t stream) {
  B200_CHECK_ARG(w_out && packed, "w4a16_dequant: null pointer");
  int rc = check_w4_shape("w4a16_dequant", K, N, group_size);
  if (rc != B200_OK) return rc;
  dim3 grid((unsigned)(K / 128), (unsigned)(N / 128));
  w4_dequant_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(w_out), static_cast<const uint8_t*>(packed), K, N,
      w4_geff(group_size));
  B200_LAUNCH_OK("w4a16_dequant");
  return B200_OK;
}

int b200_debug_w4a16_plan(int64_t N, int64_t K, int ctas, int nsub, int32_t* plan_out,
                          int32_t* first_owner_out, int32_t* contrib_out) {
  B200_CHECK_ARG(N > 0 && K > 0 && N % 128 == 0 && K % 128 == 0 && ctas >= 1 && (nsub == 1 || nsub == 2) &&
                     (N / 128) % nsub == 0 && plan_out,
                 "debug_w4a16_plan: bad arguments");
  const W4Plan pl = w4_make_plan(N, K, nsub == 2 ? 1 : 0, ctas);
  plan_out[0] = pl.units;
  plan_out[1] = pl.P;
  plan_out[2] = pl.KT;
  plan_out[3] = pl.NT;
  plan_out[4] = pl.slots;
  for (int nt = 0; nt < pl.NT; ++nt) {
    if (first_owner_out) first_owner_out[nt] = w4_first_owner(pl, nt);
    if (contrib_out) contrib_out[nt] = w4_contrib(pl, nt);
  }
  return B200_OK;
}

void b200_debug_set_trace(void* device_buffer) {
  g_w4_trace = static_cast<long long*>(device_buffer);
}

int b200_w4a16_splitk_splits(int64_t M, int64_t N, int64_t K) {
  (void)M;
  if (N <= 0 || K <= 0 || N % 128 || K % 128) return 0;
  return w4_get_plan(N, K, M).slots;
}

int b200_w4a16_gemm_splitk(float* partials, const void* A, const void* packed, int64_t M,
                           int64_t N, int64_t K, int64_t lda, int group_size, int splits,
                           b200_stream_t stream) {
  B200_CHECK_ARG(partials && A && packed, "w4a16_gemm_splitk: null pointer");
  int rc = check_w4_shape("w4a16_gemm_splitk", K, N, group_size);
  if (rc != B200_OK) return rc;
  B200_CHECK_ARG(M > 0 && M <= 128 && lda >= K, "w4a16_gemm_splitk: 1 <= M <= 128 required");
  B200_CHECK_ARG(is_aligned(A, 16) && lda % 8 == 0 && is_aligned(packed, 16) &&
                     is_aligned(partials, 16),
                 "w4a16_gemm_splitk: pointers must be 16-byte aligned, lda %% 8 == 0");
  const W4Plan plan = w4_get_plan(N, K, M);
  B200_CHECK_ARG(splits == plan.slots,
                 "w4a16_gemm_splitk: partials must have b200_w4a16_splitk_splits() = %d slots, got %d",
                 plan.slots, splits);
  return run_partial_gemm(partials, A, packed, M, N, K, lda, group_size, plan,
                          static_cast<cudaStream_t>(stream));
}

int b200_w4a16_reduce_partials(void* C, const float* partials, int splits, int64_t gemm_k,
                               const void* bias, int64_t M, int64_t N, int64_t ldc,
                               b200_stream_t stream) {
  B200_CHECK_ARG(C && partials, "w4a16_reduce_partials: null pointer");
  B200_CHECK_ARG(M >= 0 && N > 0 && N % 128 == 0 && gemm_k > 0 && gemm_k % 128 == 0 && ldc >= N,
                 "w4a16_reduce_partials: bad shape");
  const W4Plan plan = w4_get_plan(N, gemm_k, M);
  B200_CHECK_ARG(splits == plan.slots, "w4a16_reduce_partials: expected %d partial slots, got %d",
                 plan.slots, splits);
  if (M == 0) return B200_OK;
  const int64_t nvec = M * (N / 8);
  B200_PDL_LAUNCH_L(1, "w4a16_reduce", w4_reduce_kernel, (unsigned)((nvec + 255) / 256), 256, 0,
                  static_cast<cudaStream_t>(stream), static_cast<__nv_bfloat16*>(C), partials,
                  static_cast<const __nv_bfloat16*>(bias), (int)M, (int)N, ldc, M * N, plan);
  return B200_OK;
}

int64_t b200_w4a16_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 || K % 128) return 256;
  const int64_t mc = M < 128 ? M : 128, rem = M > 128 ? M % 128 : 0;  // chunk sizes the GEMM uses
  int64_t rows = (int64_t)w4_get_plan(N, K, mc).slots * mc;
  if (rem > 0) rows = std::max(rows, (int64_t)w4_get_plan(N, K, rem).slots * rem);
  return rows * N * (int64_t)sizeof(float) + 256;
}

int b200_w4a16_gemm(void* C, const void* A, const void* packed, const void* bias, int64_t M,
                    int64_t N, int64_t K, int64_t lda, int64_t ldc, int group_size,
                    void* workspace, int64_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(C && A && packed, "w4a16_gemm: null pointer");
  int rc = check_w4_shape("w4a16_gemm", K, N, group_size);
  if (rc != B200_OK) return rc;
  B200_CHECK_ARG(M >= 0 && lda >= K && ldc >= N, "w4a16_gemm: bad M/lda/ldc");
  if (M == 0) return B200_OK;
  auto st = static_cast<cudaStream_t>(stream);
  const int geff = w4_geff(group_size);

  const char* impl = getenv("B200_W4A16_IMPL");
  if (impl && impl[0] == 's') {  // bring-up path, never the default
    dim3 grid((unsigned)(N / 128), (unsigned)((M + 7) / 8));
    w4_gemm_simt_kernel<<<grid, 128, 0, st>>>(
        static_cast<__nv_bfloat16*>(C), static_cast<const __nv_bfloat16*>(A),
        static_cast<const uint8_t*>(packed), static_cast<const __nv_bfloat16*>(bias), (int)M,
        (int)N, (int)K, lda, ldc, geff);
    B200_LAUNCH_OK("w4a16_gemm_simt");
    return B200_OK;
  }

  B200_CHECK_ARG(is_aligned(A, 16) && lda % 8 == 0 && is_aligned(packed, 16),
                 "w4a16_gemm: A / packed must be 16-byte aligned, lda %% 8 == 0");
  B200_CHECK_ARG(workspace && is_aligned(workspace, 16), "w4a16_gemm: workspace required");
  // rows beyond 128 are processed in chunks of 128 (weights re-streamed; the decode path never
  // takes more than one chunk); chunks reuse the workspace in stream order
  for (int64_t m0 = 0; m0 < M; m0 += 128) {
    const int64_t mc = (M - m0) < 128 ? (M - m0) : 128;
    const W4Plan plan = w4_get_plan(N, K, mc);
    const int64_t need = (int64_t)plan.slots * mc * N * (int64_t)sizeof(float);
    if (workspace_bytes < need)
      return set_error(B200_ERR_WORKSPACE, "w4a16_gemm: workspace %lld B < required %lld B",
                       (long long)workspace_bytes, (long long)need);
    float* partials = static_cast<float*>(workspace);
    rc = run_partial_gemm(partials, static_cast<const __nv_bfloat16*>(A) + m0 * lda, packed, mc, N,
                          K, lda, group_size, plan, st);
    if (rc != B200_OK) return rc;
    const int64_t nvec = mc * (N / 8);
    B200_PDL_LAUNCH_L(1, "w4a16_reduce", w4_reduce_kernel, (unsigned)((nvec + 255) / 256), 256, 0, st,
                    static_cast<__nv_bfloat16*>(C) + m0 * ldc, (const float*)partials,
                    static_cast<const __nv_bfloat16*>(bias), (int)mc, (int)N, ldc, mc * N, plan);
  }
  return B200_OK;
}

}  // extern "C"
