// paged_attn.cu — paged-KV decode attention for sm_100a.
//
// Replaces llm::paged_kv_varlen_mha (src/kernels/attention/attn_api.cpp:14-73) for
// decode-shaped batches (q_len * group <= a few 16-row blocks).  HBM-bound: AI = group size
// FLOP/B, so tcgen05 (M >= 64 tiles) would be >= 94 % padding — the design goal is to keep
// ~100 KB of KV per SM in flight and stream every byte exactly once.
//
// Three kernels share the host side, the parameter block and the parity tests; B200_ATTN_IMPL
// selects ("stream" is the default, the other two are the bring-up / fallback paths):
//   * paged_attn_persist_kernel ("stream", default): one warp per CTA, 7 CTAs per SM.  The work
//     is the padded tile stream — the concatenation over (sequence, 16-row block of packed
//     (q token, head-in-group) rows, kv head) of ntm 16-slot KV tiles — cut into equal contiguous
//     shares, one per warp (stream-K for attention: no atomics, no wave tail).  KV tiles arrive by
//     TMA through a 4-D tensor map {64, D/64, n_kv_heads, n_slots} with SWIZZLE_128B into a
//     per-warp 3-stage ring that runs continuously across the pieces of a share; S = Q K^T and
//     O += P V use mma.sync.m16n8k16 (P rounded to the element type first, like the reference);
//     the metadata of following pieces (lengths -> ranges -> block-table window by cp.async) is
//     software-pipelined off the critical path.  A piece writes an fp32 partial O + LSE.
//   * paged_attn_mma_kernel ("mma"): the same math with one CTA per (split, row block x kv head,
//     sequence) work item and a fixed split count.
//   * paged_attn_decode_kernel ("simt"): CUDA cores only (3-D tensor map, warp-shuffle dot
//     products and online softmax); covers every dtype (incl. fp32) and head_dim.
//   paged_attn_combine_kernel merges the split partials (LSE-weighted, fixed order).
//
// Slot lookup is the reference's: block_table[block_cu_lens[b] + (pos >> log2 bs)]
// + (pos & (bs-1)), where block_table holds first-slot ids
// (src/kernels/attention/kernel/sm80_kernel_mha.cuh:148-152).
// Mask semantics follow src/kernels/attention/common/mask.h:51-86:
//   causal: kv_pos <= q_pos, q_pos = kv_len - q_len + qi
//   local : q_pos - kv_pos <= sliding_window   (window >= 0)
//   alibi : score += slope[h] * kv_pos          (after scale / soft-cap)
//   softcap: score = cap * tanh(score * sm_scale / cap)  (mha_params.h:51-66)

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

// [attn-emu:params begin]  (tools/attn_emu.cpp compiles these definitions and the whole stream
// kernel for the host)
constexpr int ATT_WARPS = 4;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int ATT_TILE = 16;         // kv slots per TMA stage
constexpr int ATT_MAX_TPS = 64;      // tiles per split (bounds the smem block table)
constexpr int ATT_TBL = ATT_MAX_TPS * ATT_TILE + 8;

struct AttnParams {
  const void* q;
  void* out;
  const int32_t* q_cu_lens;
  const int32_t* kv_cu_lens;
  const int32_t* block_table;
  const int32_t* block_cu_lens;
  const float* alibi;
  float* ws_o;    // [batch*max_q_len, n_heads, n_splits, D]
  float* ws_lse;  // [batch*max_q_len, n_heads, n_splits]
  int64_t q_stride_t, q_stride_h, o_stride_t, o_stride_h;
  int n_heads, n_kv_heads, group, n_hg;
  int block_shift, block_mask, box_rows, boxes_per_tile;
  int max_q_len, n_splits, tiles_per_split;
  int n_rb;           // mma kernel: 16-row blocks of the packed (q token, group) rows
  int* work_counter;  // (unused by the stream kernel; kept for ABI of the params block)
  int ntm;            // stream kernel: padded tiles per (sequence, row block, kv head)
  int tpw;            // stream kernel: tiles per warp (static partition of the padded tile stream)
  int stream;         // 1: split index = piece of the static stream partition (combine mirrors it)
  float scale_log2;   // sm_scale * log2(e)            (soft_cap == 0)
  float cap_in;       // sm_scale / soft_cap           (soft_cap  > 0)
  float cap_out_log2; // soft_cap * log2(e)
  int use_cap;
  int window;
};

template <int D>
struct AttnCfg {
  static constexpr int STAGES = D >= 256 ? 2 : 3;
  static constexpr int CPL = (D + 63) / 64;    // 16-byte chunks per lane per row (the last one may
                                               // be partly outside the row: head_dim 32 / 96)
  static constexpr int EPL = CPL * 8;          // elements per lane per row
  // head_dim 32 / 96: a TMA box must land on a 128-byte boundary, so the boxes of a tile sit at a
  // pitch rounded up to 64 elements (matters when a box is 1 row: block_size 1); the tile buffer
  // is sized for the padded row like the reference's 64 / 128 tiles
  static constexpr int TILE_ELEMS = ATT_TILE * ((D + 63) / 64 * 64);
};
// [attn-emu:params end]

// single-warp CTAs take at most ATT_TPS_W1 tiles, so their block-table window is small
constexpr int ATT_TPS_W1 = 32;
__host__ __device__ constexpr int att_tbl_entries(int warps) {
  return warps == 1 ? ATT_TPS_W1 * ATT_TILE + 8 : ATT_TBL;
}
template <typename T, int D, int W>
constexpr size_t attn_mma_smem_bytes() {
  return (size_t)W * AttnCfg<D>::STAGES * 2 * ATT_TILE * D * sizeof(T) +
         att_tbl_entries(W) * sizeof(int32_t) + W * AttnCfg<D>::STAGES * sizeof(uint64_t) + 128;
}

template <typename T, int D>
constexpr size_t attn_smem_bytes() {
  return (size_t)ATT_WARPS * AttnCfg<D>::STAGES * 2 * AttnCfg<D>::TILE_ELEMS * sizeof(T)  // K+V stages
         + ATT_TBL * sizeof(int32_t) + ATT_WARPS * AttnCfg<D>::STAGES * sizeof(uint64_t) + 128;
}

// [attn-emu:simt begin]
template <typename T, int D, int R>
__global__ void __launch_bounds__(ATT_THREADS, (D <= 128 ? 2 : 1))
paged_attn_decode_kernel(const __grid_constant__ CUtensorMap kmap,
                         const __grid_constant__ CUtensorMap vmap, const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  using Cfg = AttnCfg<D>;
  constexpr int STAGES = Cfg::STAGES, CPL = Cfg::CPL, EPL = Cfg::EPL;
  // head_dim 32 / 96 (the reference pads them into its 64 / 128 tiles, static_dispatch.h:16-46):
  // the row's last 64-element chunk group is only partly there; chunk (j + 8 c) exists iff
  // (j + 8 c) * 8 < D.  For head_dim % 64 == 0 the predicate is constant true.
  constexpr bool PARTIAL = (D % 64) != 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  T* stage_base = reinterpret_cast<T*>(smem_raw);
  int32_t* tbl = reinterpret_cast<int32_t*>(smem_raw + (size_t)ATT_WARPS * STAGES * 2 *
                                                           Cfg::TILE_ELEMS * sizeof(T));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tbl + ATT_TBL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  auto chunk_ok = [&](int c) { return !PARTIAL || (j + 8 * c) * 8 < D; };
  // element offset of tile row r in a stage: rows are dense inside a TMA box, boxes start on
  // 128-byte boundaries (dense for head_dim % 64 == 0)
  const int box_pitch = PARTIAL ? ((p.box_rows * D + 63) & ~63) : p.box_rows * D;
  auto row_off = [&](int r) {
    return PARTIAL ? (r / p.box_rows) * box_pitch + (r % p.box_rows) * D : r * D;
  };
  const int split = blockIdx.x;
  const int kvh = blockIdx.y / p.n_hg, hg = blockIdx.y % p.n_hg;
  const int b = blockIdx.z / p.max_q_len, qi = blockIdx.z % p.max_q_len;

  const int q_begin = p.q_cu_lens[b];
  const int q_len = p.q_cu_lens[b + 1] - q_begin;
  if (qi >= q_len) return;
  const int kv_len = p.kv_cu_lens[b + 1] - p.kv_cu_lens[b];
  const int q_pos = kv_len - q_len + qi;
  const int kv_end = q_pos + 1;
  const int kv_begin = p.window >= 0 ? max(0, q_pos - p.window) : 0;
  const int64_t tok = q_begin + qi;
  const int head0 = kvh * p.group + hg * R;
  const int rows_valid = min(R, p.group - hg * R);

  // tile range of this split, clipped to the unmasked kv range
  int t0 = split * p.tiles_per_split, t1 = t0 + p.tiles_per_split;
  t0 = max(t0, kv_begin / ATT_TILE);
  t1 = min(t1, (kv_end + ATT_TILE - 1) / ATT_TILE);

  const int64_t ws_row = (int64_t)blockIdx.z * p.n_heads;
  if (t0 >= t1) {  // nothing to attend to in this split
    if (p.n_splits > 1 && threadIdx.x < rows_valid)
      p.ws_lse[(ws_row + head0 + threadIdx.x) * p.n_splits + split] = -INFINITY;
    return;
  }

  // ---- stage the block-table window for this split --------------------------
  const int blk_cu = p.block_cu_lens[b];
  const int blk_first = (t0 * ATT_TILE) >> p.block_shift;
  const int blk_last = (min(t1 * ATT_TILE, kv_end) - 1) >> p.block_shift;
  for (int i = threadIdx.x; i <= blk_last - blk_first; i += ATT_THREADS)
    tbl[i] = p.block_table[blk_cu + blk_first + i];
  if (threadIdx.x < ATT_WARPS * STAGES) mbar_init(&bars[threadIdx.x], 1);
  fence_mbar_init();
  __syncthreads();

  T* my_stage = stage_base + (size_t)warp * STAGES * 2 * Cfg::TILE_ELEMS;
  uint64_t* my_bars = bars + warp * STAGES;
  const int n_my = (t1 - t0 - warp + ATT_WARPS - 1) / ATT_WARPS;  // tiles t0+warp, +W, ...

  auto issue = [&](int i) {  // lane 0 only
    const int tile = t0 + warp + i * ATT_WARPS;
    const int s = i % STAGES;
    T* ks = my_stage + (size_t)s * 2 * Cfg::TILE_ELEMS;
    T* vs = ks + Cfg::TILE_ELEMS;
    const int pos0 = tile * ATT_TILE;
    int nbox = 0;
    for (int bx = 0; bx < p.boxes_per_tile; ++bx)
      if (pos0 + bx * p.box_rows < kv_end) ++nbox;
    mbar_arrive_expect_tx(&my_bars[s], (uint32_t)(nbox * 2 * p.box_rows * D * sizeof(T)));
    for (int bx = 0; bx < nbox; ++bx) {
      const int pos = pos0 + bx * p.box_rows;
      const int slot0 = tbl[(pos >> p.block_shift) - blk_first] + (pos & p.block_mask);
      tma_load_3d(ks + bx * box_pitch, &kmap, &my_bars[s], 0, kvh, slot0);
      tma_load_3d(vs + bx * box_pitch, &vmap, &my_bars[s], 0, kvh, slot0);
    }
  };

  if (lane == 0) {
    if (warp == 0) {
      prefetch_tensormap(&kmap);
      prefetch_tensormap(&vmap);
    }
    for (int i = 0; i < STAGES && i < n_my; ++i) issue(i);
  }

  // ---- query rows of this head group -> registers (fp32) --------------------
  float qf[R][EPL];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int h = head0 + (r < rows_valid ? r : 0);
    const T* qrow = static_cast<const T*>(p.q) + tok * p.q_stride_t + (int64_t)h * p.q_stride_h;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      uint4 raw = chunk_ok(c) ? ld_v4(qrow + (j + 8 * c) * 8) : make_uint4(0, 0, 0, 0);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = Num<T>::unpack(w[e]);
        qf[r][c * 8 + 2 * e] = f.x;
        qf[r][c * 8 + 2 * e + 1] = f.y;
      }
    }
  }
  float slope_log2[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    slope_log2[r] = p.alibi ? p.alibi[head0 + (r < rows_valid ? r : 0)] * 1.4426950408889634f : 0.f;

  float m[R], l[R], acc[R][EPL];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[r][e] = 0.f;
  }

  // ---- main loop: this warp's tiles ----------------------------------------
  for (int i = 0; i < n_my; ++i) {
    const int s = i % STAGES;
    const uint32_t phase = (i / STAGES) & 1;
    const T* ks = my_stage + (size_t)s * 2 * Cfg::TILE_ELEMS;
    const T* vs = ks + Cfg::TILE_ELEMS;
    const int pos0 = (t0 + warp + i * ATT_WARPS) * ATT_TILE;
    mbar_wait(&my_bars[s], phase);

    // scores: lane group g handles slots g, g+4, g+8, g+12 of the tile
    float sc[4][R];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const T* krow = ks + row_off(it * 4 + g);
      float kf[EPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        uint4 raw = chunk_ok(c) ? ld_v4(krow + (j + 8 * c) * 8) : make_uint4(0, 0, 0, 0);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = Num<T>::unpack(w[e]);
          kf[c * 8 + 2 * e] = f.x;
          kf[c * 8 + 2 * e + 1] = f.y;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) a = fmaf(qf[r][e], kf[e], a);
        sc[it][r] = a;
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
#pragma unroll
      for (int it = 0; it < 4; ++it)
#pragma unroll
        for (int r = 0; r < R; ++r) sc[it][r] += __shfl_xor_sync(0xffffffffu, sc[it][r], o);

    // online softmax (exp2 domain), private to the lane group
    bool valid[4];
    float pr[4][R];
    bool need_rescale = false;
    float corr[R];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int pos = pos0 + it * 4 + g;
      valid[it] = pos >= kv_begin && pos < kv_end;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float x[4];
      float mx = m[r];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int pos = pos0 + it * 4 + g;
        float v = p.use_cap ? tanhf(sc[it][r] * p.cap_in) * p.cap_out_log2 : sc[it][r] * p.scale_log2;
        v = fmaf(slope_log2[r], (float)pos, v);
        x[it] = valid[it] ? v : -INFINITY;
        mx = fmaxf(mx, x[it]);
      }
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      corr[r] = exp2f(m[r] - ms);
      float sum = 0.f;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        pr[it][r] = exp2f(x[it] - ms);
        sum += pr[it][r];
      }
      l[r] = fmaf(l[r], corr[r], sum);
      m[r] = mx;
      need_rescale |= (corr[r] != 1.f);
    }
    if (__any_sync(0xffffffffu, need_rescale)) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[r][e] *= corr[r];
    }

    // acc += P V
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (valid[it]) {  // masked rows may hold stale bytes (NaN) — never touch them
        const T* vrow = vs + row_off(it * 4 + g);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          uint4 raw = chunk_ok(c) ? ld_v4(vrow + (j + 8 * c) * 8) : make_uint4(0, 0, 0, 0);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = Num<T>::unpack(w[e]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
              acc[r][c * 8 + 2 * e] = fmaf(pr[it][r], f.x, acc[r][c * 8 + 2 * e]);
              acc[r][c * 8 + 2 * e + 1] = fmaf(pr[it][r], f.y, acc[r][c * 8 + 2 * e + 1]);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0 && i + STAGES < n_my) issue(i + STAGES);
  }

  // ---- merge the 4 lane groups of the warp ----------------------------------
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[r], off);
      const float lo = __shfl_xor_sync(0xffffffffu, l[r], off);
      const float mn = fmaxf(m[r], mo);
      const float ms = (mn == -INFINITY) ? 0.f : mn;
      const float a = exp2f(m[r] - ms), bb = exp2f(mo - ms);
      l[r] = l[r] * a + lo * bb;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[r][e], off);
        acc[r][e] = acc[r][e] * a + ao * bb;
      }
      m[r] = mn;
    }
  }

  // ---- merge warps through shared memory (each warp reuses its own stages) ---
  float* red = reinterpret_cast<float*>(my_stage);  // [R][D] acc, then [R] m, [R] l
  if (g == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        if (!chunk_ok(c)) continue;
#pragma unroll
        for (int e = 0; e < 8; ++e) red[r * D + (j + 8 * c) * 8 + e] = acc[r][c * 8 + e];
      }
      if (j == 0) {
        red[R * D + r] = m[r];
        red[R * D + R + r] = l[r];
      }
    }
  }
  __syncthreads();
  constexpr size_t WARP_STRIDE = (size_t)STAGES * 2 * Cfg::TILE_ELEMS * sizeof(T) / sizeof(float);
  const float* red0 = reinterpret_cast<const float*>(stage_base);
  for (int idx = threadIdx.x; idx < rows_valid * D; idx += ATT_THREADS) {
    const int r = idx / D, d = idx % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) M = fmaxf(M, red0[w * WARP_STRIDE + R * D + r]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) {
      const float sc_w = exp2f(red0[w * WARP_STRIDE + R * D + r] - M);
      L = fmaf(red0[w * WARP_STRIDE + R * D + R + r], sc_w, L);
      O = fmaf(red0[w * WARP_STRIDE + r * D + d], sc_w, O);
    }
    const float o = O / L;
    const int h = head0 + r;
    if (p.n_splits == 1) {
      static_cast<T*>(p.out)[tok * p.o_stride_t + (int64_t)h * p.o_stride_h + d] = Num<T>::from_f(o);
    } else {
      p.ws_o[((ws_row + h) * p.n_splits + split) * D + d] = o;
      if (d == 0) p.ws_lse[(ws_row + h) * p.n_splits + split] = M + log2f(L);
    }
  }
}
// [attn-emu:simt end]


// ===========================================================================
// Tensor-core variant (legacy HMMA path; M = 16 rows is the right tile for decode: tcgen05's
// M >= 64 tiles would waste >= 94 % of the MMA on G*q_len = 4 rows, exactly the reference's
// problem with its 64-row tile).  Rows of the MMA tile are the packed (q token, group) pairs of
// one kv head, like the reference packs q_len x group into M (sm80_kernel_mha.cuh:208-262), so
// speculative / multi-token decode (q_len * G <= 16) costs one KV pass.  S = Q K^T and O += P V
// use mma.sync.m16n8k16 with fp32 accumulation; P is cast to the element type before PV like the
// reference (sm80_collective_mha.cuh:289-290).  K/V tiles land in shared memory through a 4-D
// tensor map {64, D/64, n_kv_heads, n_slots} with SWIZZLE_128B so ldmatrix is (nearly)
// conflict free; the instruction count per 16-slot tile drops ~7x vs the CUDA-core kernel,
// which leaves the warp schedulers to the memory pipeline.
// ===========================================================================
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4],
                                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4],
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// transpose of an 8x8 b16 matrix held one 32-bit register per lane (row lane / 4, columns
// 2 (lane % 4), +1): turns an accumulator-layout pair into a B-operand-layout pair
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// 128-byte swizzle as the TMA applies it: 16-byte chunk index ^= (address bits [7,10))
__device__ __forceinline__ uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }

template <typename T, int D, int W>
__global__ void __launch_bounds__(W * 32, (W == 1 ? (D <= 128 ? 6 : 2) : (D <= 128 ? 2 : 1)))
paged_attn_mma_kernel(const __grid_constant__ CUtensorMap kmap,
                      const __grid_constant__ CUtensorMap vmap, const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  using Cfg = AttnCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int KS = D / 16;   // k-steps of S = Q K^T
  constexpr int NB = D / 8;    // n-blocks of O
  constexpr int ROWB = D * (int)sizeof(T);  // bytes per slot row in a tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  T* stage_base = reinterpret_cast<T*>(smem_raw);
  int32_t* tbl = reinterpret_cast<int32_t*>(smem_raw + (size_t)W * STAGES * 2 *
                                                           Cfg::TILE_ELEMS * sizeof(T));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tbl + att_tbl_entries(W));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, kvh = blockIdx.y;
  const int b = blockIdx.z / p.n_rb, rb = blockIdx.z % p.n_rb;
  const int G = p.group;

  const int q_begin = p.q_cu_lens[b];
  const int q_len = p.q_cu_lens[b + 1] - q_begin;
  const int rows_total = q_len * G;
  const int row0 = rb * 16;
  if (row0 >= rows_total) return;
  const int n_rows = min(16, rows_total - row0);
  const int kv_len = p.kv_cu_lens[b + 1] - p.kv_cu_lens[b];
  const int q_pos0 = kv_len - q_len;  // position of query token 0

  // rows owned by this lane in the C / A fragments: r_lo = lane/4, r_hi = r_lo + 8
  int row_qi[2], row_head[2], row_end[2], row_begin[2];
  bool row_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (lane >> 2) + 8 * h;
    row_ok[h] = r < n_rows;
    const int row = row0 + (row_ok[h] ? r : 0);
    row_qi[h] = row / G;
    row_head[h] = kvh * G + (row - row_qi[h] * G);
    const int qp = q_pos0 + row_qi[h];
    row_end[h] = row_ok[h] ? qp + 1 : 0;          // invalid rows attend to nothing
    row_begin[h] = p.window >= 0 ? max(0, qp - p.window) : 0;
  }
  const int qi_min = row0 / G, qi_max = (row0 + n_rows - 1) / G;
  const int kv_end = q_pos0 + qi_max + 1;
  const int kv_begin = p.window >= 0 ? max(0, q_pos0 + qi_min - p.window) : 0;

  int t0 = split * p.tiles_per_split, t1 = t0 + p.tiles_per_split;
  t0 = max(t0, kv_begin / ATT_TILE);
  t1 = min(t1, (kv_end + ATT_TILE - 1) / ATT_TILE);

  if (t0 >= t1) {
    if (p.n_splits > 1 && threadIdx.x < n_rows) {
      const int row = row0 + threadIdx.x, qi = row / G, head = kvh * G + (row - qi * G);
      p.ws_lse[(((int64_t)b * p.max_q_len + qi) * p.n_heads + head) * p.n_splits + split] =
          -INFINITY;
    }
    return;
  }

  const int blk_cu = p.block_cu_lens[b];
  const int blk_first = (t0 * ATT_TILE) >> p.block_shift;
  const int blk_last = (min(t1 * ATT_TILE, kv_end) - 1) >> p.block_shift;
  for (int i = threadIdx.x; i <= blk_last - blk_first; i += (W * 32))
    tbl[i] = p.block_table[blk_cu + blk_first + i];
  if (threadIdx.x < W * STAGES) mbar_init(&bars[threadIdx.x], 1);
  fence_mbar_init();
  __syncthreads();

  T* my_stage = stage_base + (size_t)warp * STAGES * 2 * Cfg::TILE_ELEMS;
  uint64_t* my_bars = bars + warp * STAGES;
  const int n_my = (t1 - t0 - warp + W - 1) / W;

  auto issue = [&](int i) {  // lane 0 only
    const int tile = t0 + warp + i * W;
    const int s = i % STAGES;
    T* ks = my_stage + (size_t)s * 2 * Cfg::TILE_ELEMS;
    T* vs = ks + Cfg::TILE_ELEMS;
    const int pos0 = tile * ATT_TILE;
    int nbox = 0;
    for (int bx = 0; bx < p.boxes_per_tile; ++bx)
      if (pos0 + bx * p.box_rows < kv_end) ++nbox;
    mbar_arrive_expect_tx(&my_bars[s], (uint32_t)(nbox * 2 * p.box_rows * D * sizeof(T)));
    for (int bx = 0; bx < nbox; ++bx) {
      const int pos = pos0 + bx * p.box_rows;
      const int slot0 = tbl[(pos >> p.block_shift) - blk_first] + (pos & p.block_mask);
      tma_load_4d(ks + bx * p.box_rows * D, &kmap, &my_bars[s], 0, 0, kvh, slot0);
      tma_load_4d(vs + bx * p.box_rows * D, &vmap, &my_bars[s], 0, 0, kvh, slot0);
    }
  };
  if (lane == 0) {
    if (warp == 0) {
      prefetch_tensormap(&kmap);
      prefetch_tensormap(&vmap);
    }
    for (int i = 0; i < STAGES && i < n_my; ++i) issue(i);
  }

  // ---- Q as A fragments (row-major 16 x 16 per k-step) ------------------------
  uint32_t qa[KS][4];
  {
    const T* qrow[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      qrow[h] = static_cast<const T*>(p.q) + (int64_t)(q_begin + row_qi[h]) * p.q_stride_t +
                (int64_t)row_head[h] * p.q_stride_h + (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        qa[ks][h] = row_ok[h] ? *reinterpret_cast<const uint32_t*>(qrow[h] + ks * 16) : 0u;
        qa[ks][2 + h] = row_ok[h] ? *reinterpret_cast<const uint32_t*>(qrow[h] + ks * 16 + 8) : 0u;
      }
    }
  }
  float slope_log2[2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
    slope_log2[h] = p.alibi ? p.alibi[row_head[h]] * 1.4426950408889634f : 0.f;

  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nb][e] = 0.f;

  // per-lane ldmatrix row addressing (see the fragment notes above)
  const int lm = lane >> 3, lr = lane & 7;

  for (int i = 0; i < n_my; ++i) {
    const int s = i % STAGES;
    const uint32_t phase = (i / STAGES) & 1;
    T* ks_t = my_stage + (size_t)s * 2 * Cfg::TILE_ELEMS;
    T* vs_t = ks_t + Cfg::TILE_ELEMS;
    const uint32_t k_base = smem_u32(ks_t), v_base = smem_u32(vs_t);
    const int pos0 = (t0 + warp + i * W) * ATT_TILE;
    mbar_wait(&my_bars[s], phase);

    // slots at or beyond kv_end may hold stale shared memory or another owner's data (possibly
    // NaN/Inf): P is 0 there but 0 * NaN would poison the PV MMA, so zero those V rows.
    const bool boundary = pos0 + ATT_TILE > kv_end;
    if (boundary) {
      const int first_bad = kv_end - pos0;  // 1..15
      for (int c = lane; c < (ATT_TILE - first_bad) * (ROWB / 16); c += 32) {
        const int row = first_bad + c / (ROWB / 16), ch = c % (ROWB / 16);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(vs_t) + row * ROWB + ch * 16) =
            make_uint4(0, 0, 0, 0);  // zeroing a whole row: the swizzle permutes within the row
      }
      __syncwarp();
    }

    // ---- S = Q K^T : two n-blocks of 8 slots ------------------------------------
    float sacc[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[nb][e] = 0.f;
#pragma unroll
      for (int kq = 0; kq < KS / 2; ++kq) {  // one ldmatrix.x4 = 8 slots x 32 d = two k-steps
        uint32_t bf[4];
        const uint32_t off = (uint32_t)((nb * 8 + lr) * ROWB + (kq * 32 + lm * 8) * (int)sizeof(T));
        ldsm_x4(bf, swz128(k_base + off));
        mma_16816<T>(sacc[nb], qa[2 * kq], bf[0], bf[1]);
        mma_16816<T>(sacc[nb], qa[2 * kq + 1], bf[2], bf[3]);
      }
    }

    // ---- mask + online softmax (exp2 domain) ------------------------------------
    float corr[2];
    bool need_rescale = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float x[4];
      float mx = m[h];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int pos = pos0 + nb * 8 + (lane & 3) * 2 + e;
          const float sc = sacc[nb][2 * h + e];
          float v = p.use_cap ? tanhf(sc * p.cap_in) * p.cap_out_log2 : sc * p.scale_log2;
          v = fmaf(slope_log2[h], (float)pos, v);
          const bool ok = pos >= row_begin[h] && pos < row_end[h];
          x[nb * 2 + e] = ok ? v : -INFINITY;
          mx = fmaxf(mx, x[nb * 2 + e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      corr[h] = (mx == m[h]) ? 1.f : exp2f(m[h] - ms);  // unchanged max (or still -inf): no rescale
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x[c] = exp2f(x[c] - ms);
        sum += x[c];
      }
      l[h] = fmaf(l[h], corr[h], sum);  // per-lane partial; quad-reduced after the loop
      m[h] = mx;
      need_rescale |= (corr[h] != 1.f);
      sacc[0][2 * h] = x[0];
      sacc[0][2 * h + 1] = x[1];
      sacc[1][2 * h] = x[2];
      sacc[1][2 * h + 1] = x[3];
    }
    if (__any_sync(0xffffffffu, need_rescale)) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        o[nb][0] *= corr[0];
        o[nb][1] *= corr[0];
        o[nb][2] *= corr[1];
        o[nb][3] *= corr[1];
      }
    }
    // P (16 rows x 16 slots) as the A fragment of the PV MMA, cast to T
    uint32_t pa[4];
    pa[0] = Num<T>::pack(sacc[0][0], sacc[0][1]);
    pa[1] = Num<T>::pack(sacc[0][2], sacc[0][3]);
    pa[2] = Num<T>::pack(sacc[1][0], sacc[1][1]);
    pa[3] = Num<T>::pack(sacc[1][2], sacc[1][3]);

    // ---- O += P V : ldmatrix.trans gives V^T fragments, 16 d per instruction ------
#pragma unroll
    for (int dq = 0; dq < NB / 2; ++dq) {
      uint32_t bf[4];
      const uint32_t off =
          (uint32_t)(((lm & 1) * 8 + lr) * ROWB + (dq * 16 + (lm >> 1) * 8) * (int)sizeof(T));
      ldsm_x4_trans(bf, swz128(v_base + off));
      mma_16816<T>(o[2 * dq], pa, bf[0], bf[1]);
      mma_16816<T>(o[2 * dq + 1], pa, bf[2], bf[3]);
    }
    __syncwarp();
    if (i + STAGES < n_my) {
      if (boundary) fence_proxy_async_smem();  // our generic zero-stores before the next TMA write
      if (lane == 0) issue(i + STAGES);
    }
  }

  // ---- finish the row sums across the quad ---------------------------------------
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l[h] += __shfl_xor_sync(0xffffffffu, l[h], 1);
    l[h] += __shfl_xor_sync(0xffffffffu, l[h], 2);
  }

  // ---- merge the warps through shared memory (each warp reuses its own stages) -----
  float* red = reinterpret_cast<float*>(my_stage);  // [16][D] O, then [16] m, [16] l
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (lane >> 2) + 8 * h;
    if (r < n_rows) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        *reinterpret_cast<float2*>(&red[r * D + nb * 8 + (lane & 3) * 2]) =
            make_float2(o[nb][2 * h], o[nb][2 * h + 1]);
      if ((lane & 3) == 0) {
        red[16 * D + r] = m[h];
        red[16 * D + 16 + r] = l[h];
      }
    }
  }
  __syncthreads();
  constexpr size_t WARP_STRIDE = (size_t)STAGES * 2 * Cfg::TILE_ELEMS * sizeof(T) / sizeof(float);
  const float* red0 = reinterpret_cast<const float*>(stage_base);
  for (int idx = threadIdx.x; idx < n_rows * D; idx += (W * 32)) {
    const int r = idx / D, d = idx % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < W; ++w) M = fmaxf(M, red0[w * WARP_STRIDE + 16 * D + r]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float sc_w = exp2f(red0[w * WARP_STRIDE + 16 * D + r] - M);
      L = fmaf(red0[w * WARP_STRIDE + 16 * D + 16 + r], sc_w, L);
      O = fmaf(red0[w * WARP_STRIDE + r * D + d], sc_w, O);
    }
    const float ov = O / L;
    const int row = row0 + r, qi = row / G, head = kvh * G + (row - qi * G);
    if (p.n_splits == 1) {
      static_cast<T*>(p.out)[(int64_t)(q_begin + qi) * p.o_stride_t + (int64_t)head * p.o_stride_h + d] =
          Num<T>::from_f(ov);
    } else {
      const int64_t wrow = ((int64_t)b * p.max_q_len + qi) * p.n_heads + head;
      p.ws_o[(wrow * p.n_splits + split) * D + d] = ov;
      if (d == 0) p.ws_lse[wrow * p.n_splits + split] = M + log2f(L);
    }
  }
}


// ===========================================================================
// Stream variant of the tensor-core kernel (default): one warp per CTA and a STATIC, equal
// partition of the padded tile stream — the concatenation over (sequence, row block, kv head) of
// n_tiles_max tiles each; warp w owns stream positions [w*tpw, (w+1)*tpw).  Every warp therefore
// streams the same number of KV tiles (stream-K for attention): no work counter, no wave tail,
// a handful of pieces per warp.  A piece (warp range x one sequence) is a split of that sequence;
// the warp keeps ONE continuous TMA ring across its pieces and runs the lengths -> block table
// -> shared memory prefetch of the following pieces one stage per rotation, so the dependent
// global accesses are off the critical path.  Partial O / LSE go to the workspace; the combine
// pass recomputes the same partition to know which pieces exist.
// ===========================================================================
// [attn-emu:persist begin]
constexpr int ATT_P_TPS_MAX = 32;                            // tiles per item (upper bound)
constexpr int ATT_P_TBL = ATT_P_TPS_MAX * ATT_TILE + 8;      // block-table window entries (bs = 1)
// Occupancy variant (OCC = 1, B200_ATTN_OCC=1): the warp-state profile of the default variant
// shows warps waiting on their own instruction latencies 60 % of the time and on KV data 5 %
// (profiles/r01_ncu_paged_attn_stalls.md), i.e. it is latency bound at 7 warps per SM.  Round 2
// measured the first cut of this variant (2 TMA stages, 11 CTAs per SM): 114.8 us vs 96.6 us —
// the shallower ring costs more than the extra warps give (profiles/r02_attn_variants.md).  What
// is left of it: the full 3-stage ring with a 264-entry block-table window, which fits 8 instead
// of 7 one-warp CTAs into an SM's shared memory.
constexpr int ATT_P_TBL_OCC = 256 + 8;
constexpr int ATT_OCC_CTAS = 8;
__host__ __device__ constexpr int att_p_tbl(int occ) { return occ ? ATT_P_TBL_OCC : ATT_P_TBL; }


// One work item moving through the claim pipeline.  A stage runs once per rotation, so every
// dependent global access (atomic -> lengths -> block table) has a whole item's duration to land.
struct ItemMeta {
  int valid;                       // 0 = past the end of the work list
  int b, kvh, rb, split;
  int q_begin, q_end, kv_b0, kv_b1, blk_cu;  // stage A: raw lengths
  int q_len, kv_len;
  int t0, n_tiles;                 // stage B: tile range (n_tiles = 0: nothing to attend to)
  int t_hi;                        // stage A: end of this piece's window inside its sequence
  int kv_end, kv_begin, blk_first, n_ent;
  int tbl_buf;                     // stage B: which of the 3 table buffers holds its window
};

// TR = 1 (B200_ATTN_TR=1, decode shapes with group * max_q_len <= 8 packed rows): the transposed
// tile.  S^T[16 keys x 8 rows] = K . Q^T and O^T[D x 8 rows] += V^T . P^T put the keys (and the
// head dimension) on the MMA's M and the query rows on its 8-wide N: half the HMMAs per tile, 4
// instead of 8 scores per lane in the softmax, an O accumulator of D/4 instead of D/2 registers
// and Q fragments of half the size — the instruction chain per tile and the register count are
// what bound the default variant (profiles/r01_ncu_paged_attn_stalls.md).  The fragment algebra
// is modelled lane by lane in tools/attn_tr_model.py.  Row blocks are 8 rows instead of 16; the
// plan selects TR only when all packed rows fit one block, so the work partition (and the combine
// pass that mirrors it) is the same as the default's.
template <typename T, int D, int OCC, int TR = 0>
__global__ void __launch_bounds__(32, (D <= 128 ? (OCC ? ATT_OCC_CTAS : 8) : 3))
paged_attn_persist_kernel(const __grid_constant__ CUtensorMap kmap,
                          const __grid_constant__ CUtensorMap vmap, const AttnParams p,
                          int64_t total_tiles, int n_seq) {
  // Under programmatic dependent launch this kernel becomes resident while its predecessor (the
  // RoPE + KV-write consumer) still runs.  Everything up to the first Q / KV access depends only on
  // the step's metadata (lengths, block tables: inputs of the step, older than any kernel of it), so
  // the pipeline fill below — three dependent global round trips — runs in the predecessor's
  // shadow; griddepcontrol.wait sits just before the first query load.
  pdl_launch_dependents();
  using Cfg = AttnCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int P_TBL = att_p_tbl(OCC);
  constexpr int KS = D / 16, NB = TR ? D / 16 : D / 8;  // NB: accumulator blocks of O (O^T)
  constexpr int ROWS = TR ? 8 : 16;                      // packed (q token, head) rows per block
  constexpr int QR = TR ? 2 : 4;                         // Q fragment registers per k-step
  constexpr int ROWB = D * (int)sizeof(T);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  T* stage_base = reinterpret_cast<T*>(smem_raw);
  int32_t* tbl_base = reinterpret_cast<int32_t*>(smem_raw + (size_t)STAGES * 2 * Cfg::TILE_ELEMS * sizeof(T));
  // this warp's slice of the padded tile stream
  const int64_t g0 = (int64_t)blockIdx.x * p.tpw;
  const int64_t g1 = g0 + p.tpw < total_tiles ? g0 + p.tpw : total_tiles;
  if (g0 >= g1) return;
  const int seq_first = (int)(g0 / p.ntm), seq_last = (int)((g1 - 1) / p.ntm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tbl_base + 3 * P_TBL);
  const int lane = threadIdx.x;
  const int G = p.group;

  if (lane < STAGES) mbar_init(&bars[lane], 1);
  fence_mbar_init();
  __syncwarp();
  if (lane == 0) {
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }

  // ---- claim pipeline stages --------------------------------------------------------------
  int next_seq = seq_first;
  int tbl_rot = 0;  // rotating table buffer index (3 buffers: current, next, the one being filled)
  auto stage0 = [&]() -> int { return next_seq++; };  // pieces are the sequences the slice touches
  auto stageA = [&](int seq) -> ItemMeta {  // decode + issue the length loads
    ItemMeta it;
    it.valid = seq <= seq_last;
    const int sc = it.valid ? seq : seq_first;
    it.kvh = sc % p.n_kv_heads;
    int rest = sc / p.n_kv_heads;
    it.rb = rest % p.n_rb;
    it.b = rest / p.n_rb;
    // tile window of this piece inside its sequence, and which piece of the sequence it is
    const int64_t base = (int64_t)sc * p.ntm;
    it.t0 = (int)((g0 > base ? g0 : base) - base);
    it.t_hi = (int)((g1 < base + p.ntm ? g1 : base + p.ntm) - base);
    it.split = (int)((base + it.t0) / p.tpw - base / p.tpw);
    it.q_begin = p.q_cu_lens[it.b];
    it.q_end = p.q_cu_lens[it.b + 1];
    it.kv_b0 = p.kv_cu_lens[it.b];
    it.kv_b1 = p.kv_cu_lens[it.b + 1];
    it.blk_cu = p.block_cu_lens[it.b];
    it.n_tiles = 0;
    return it;
  };
  auto stageB = [&](ItemMeta& it) {  // tile range + issue the block-table loads
    it.q_len = it.q_end - it.q_begin;
    it.kv_len = it.kv_b1 - it.kv_b0;
    const int t_lo = it.t0, t_hi = it.t_hi;
    it.n_tiles = 0;
    it.n_ent = 0;
    const int rows_total = it.q_len * G, row0 = it.rb * ROWS;
    if (!it.valid || row0 >= rows_total) {
      it.q_len = it.valid ? it.q_len : 0;
      return;
    }
    const int n_rows = min(ROWS, rows_total - row0);
    const int q_pos0 = it.kv_len - it.q_len;
    const int qi_min = row0 / G, qi_max = (row0 + n_rows - 1) / G;
    it.kv_end = q_pos0 + qi_max + 1;
    it.kv_begin = p.window >= 0 ? max(0, q_pos0 + qi_min - p.window) : 0;
    const int t0 = max(t_lo, it.kv_begin / ATT_TILE);
    const int t1 = min(t_hi, (it.kv_end + ATT_TILE - 1) / ATT_TILE);
    if (t0 >= t1) return;
    it.t0 = t0;
    it.n_tiles = t1 - t0;
    it.blk_first = (t0 * ATT_TILE) >> p.block_shift;
    it.n_ent = ((min(t1 * ATT_TILE, it.kv_end) - 1) >> p.block_shift) - it.blk_first + 1;
    // asynchronous copy of the block-table window: lands while the previous pieces are processed
    it.tbl_buf = tbl_rot;
    tbl_rot = tbl_rot == 2 ? 0 : tbl_rot + 1;
    const uint32_t dst = smem_u32(tbl_base + it.tbl_buf * P_TBL);
    const int32_t* src = p.block_table + it.blk_cu + it.blk_first;
    for (int e = lane; e < it.n_ent; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + e * 4), "l"(src + e)
                   : "memory");
  };
  auto stageC = [&]() {  // every table window requested so far has landed
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  };

  // TMA for tile `ti` of item `it` into ring slot of stream position `g`
  auto issue = [&](const ItemMeta& it, int ti, int g) {  // lane 0 only
    const int s = g % STAGES;
    T* ks = stage_base + (size_t)s * 2 * Cfg::TILE_ELEMS;
    T* vs = ks + Cfg::TILE_ELEMS;
    const int32_t* tbl = tbl_base + it.tbl_buf * P_TBL;
    const int pos0 = (it.t0 + ti) * ATT_TILE;
    int nbox = 0;
    for (int bx = 0; bx < p.boxes_per_tile; ++bx)
      if (pos0 + bx * p.box_rows < it.kv_end) ++nbox;
    mbar_arrive_expect_tx(&bars[s], (uint32_t)(nbox * 2 * p.box_rows * D * sizeof(T)));
    for (int bx = 0; bx < nbox; ++bx) {
      const int pos = pos0 + bx * p.box_rows;
      const int slot0 = tbl[(pos >> p.block_shift) - it.blk_first] + (pos & p.block_mask);
      tma_load_4d(ks + bx * p.box_rows * D, &kmap, &bars[s], 0, 0, it.kvh, slot0);
      tma_load_4d(vs + bx * p.box_rows * D, &vmap, &bars[s], 0, 0, it.kvh, slot0);
    }
  };

  // query fragments of an item (A operand of S = Q K^T), rows beyond the item's rows are zero
  auto load_q = [&](const ItemMeta& it, uint32_t (&qa)[KS][QR]) {
    const int rows_total = it.q_len * G, row0 = it.rb * ROWS;
    const int n_rows = min(ROWS, rows_total - row0);
    if constexpr (TR) {
      // [tr-emu:load_q begin]  (tools/attn_tr_emu.cpp compiles this block for the host)
      // B operand of S^T = K Q^T: lane (g, t) holds Q[row g][16 ks + 2t, +1] and [.. + 8, + 9]
      const int r = lane >> 2;
      const bool ok = it.n_tiles > 0 && r < n_rows;
      const int row = row0 + (ok ? r : 0), qi = ok ? row / G : 0;
      const int head = it.kvh * G + (ok ? row - qi * G : 0);
      const T* qrow = static_cast<const T*>(p.q) + (int64_t)(it.q_begin + qi) * p.q_stride_t +
                      (int64_t)head * p.q_stride_h + (lane & 3) * 2;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qa[ks][0] = ok ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16) : 0u;
        qa[ks][1] = ok ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8) : 0u;
      }
      // [tr-emu:load_q end]
    } else {
      // [def-emu:load_q begin]  (the default blocks run through the same host harness: they are
      // GPU-validated, so they validate the harness's ldmatrix / mma.sync emulation)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = (lane >> 2) + 8 * h;
        const bool ok = it.n_tiles > 0 && r < n_rows;
        const int row = row0 + (ok ? r : 0), qi = ok ? row / G : 0;
        const int head = it.kvh * G + (ok ? row - qi * G : 0);
        const T* qrow = static_cast<const T*>(p.q) + (int64_t)(it.q_begin + qi) * p.q_stride_t +
                        (int64_t)head * p.q_stride_h + (lane & 3) * 2;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          qa[ks][h] = ok ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16) : 0u;
          qa[ks][2 + h] = ok ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8) : 0u;
        }
      }
      // [def-emu:load_q end]
    }
  };

  // ---- fill the pipeline (only here are the dependent latencies exposed) ----------------------
  ItemMeta cur = stageA(stage0());
  stageB(cur);
  ItemMeta nxt = stageA(stage0());
  stageB(nxt);
  stageC();                         // cur's and nxt's table windows are in shared memory
  ItemMeta sB = stageA(stage0());
  stageB(sB);                       // its table window is in flight into the third buffer
  ItemMeta sA = stageA(stage0());   // lengths in flight
  int s0 = stage0();
  uint32_t qa[KS][QR], qn[KS][QR];
  pdl_wait();  // q and the newest KV slots are the predecessor's output
  load_q(cur, qa);
  load_q(nxt, qn);

  int g_cons = 0, g_iss = 0;         // stream positions (tiles consumed / issued)
  int cur_issued = 0, nxt_issued = 0;
  auto issue_ahead = [&]() {          // keep the ring full: current item first, then the next one
    while (g_iss - g_cons < STAGES) {
      if (cur_issued < cur.n_tiles) {
        if (lane == 0) issue(cur, cur_issued, g_iss);
        ++cur_issued;
      } else if (nxt.valid && nxt_issued < nxt.n_tiles) {
        if (lane == 0) issue(nxt, nxt_issued, g_iss);
        ++nxt_issued;
      } else {
        break;
      }
      ++g_iss;
    }
  };
  issue_ahead();

  const int lm = lane >> 3, lr = lane & 7;
  while (cur.valid) {
    // ---- per-item row bookkeeping ------------------------------------------------------
    const int rows_total = cur.q_len * G, row0 = cur.rb * ROWS;
    const int n_rows = min(ROWS, rows_total - row0);
    const int q_pos0 = cur.kv_len - cur.q_len;
    // the two rows a lane works for: rows g, g + 8 of the block (TR: rows 2t, 2t + 1)
    int row_end[2], row_begin[2], row_head[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = TR ? (lane & 3) * 2 + h : (lane >> 2) + 8 * h;
      const bool ok = r < n_rows;
      const int row = row0 + (ok ? r : 0), qi = row / G;
      row_head[h] = cur.kvh * G + (row - qi * G);
      const int qp = q_pos0 + qi;
      // (TR: a padding row gets the first row's bounds instead of an empty range — its Q is zero and
      // its column of O^T is never written — so that the interior-tile test below is warp-uniform)
      row_end[h] = (ok || TR) ? qp + 1 : 0;
      row_begin[h] = p.window >= 0 ? max(0, qp - p.window) : 0;
    }
    float slope_log2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      slope_log2[h] = (p.alibi && cur.n_tiles > 0) ? p.alibi[row_head[h]] * 1.4426950408889634f : 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float o[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nb][e] = 0.f;

    for (int i = 0; i < cur.n_tiles; ++i) {
      const int s = g_cons % STAGES;
      const uint32_t phase = (g_cons / STAGES) & 1;
      T* ks_t = stage_base + (size_t)s * 2 * Cfg::TILE_ELEMS;
      T* vs_t = ks_t + Cfg::TILE_ELEMS;
      const uint32_t k_base = smem_u32(ks_t), v_base = smem_u32(vs_t);
      const int pos0 = (cur.t0 + i) * ATT_TILE;
      mbar_wait(&bars[s], phase);

      const bool boundary = pos0 + ATT_TILE > cur.kv_end;
      if (boundary) {  // zero V rows past the causal end (stale / foreign data, maybe NaN)
        const int first_bad = cur.kv_end - pos0;
        for (int c = lane; c < (ATT_TILE - first_bad) * (ROWB / 16); c += 32) {
          const int row = first_bad + c / (ROWB / 16), ch = c % (ROWB / 16);
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(vs_t) + row * ROWB + ch * 16) =
              make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
      }

      if constexpr (TR) {
        // [tr-emu:tile begin]
        // S^T = K Q^T: A = the tile's 16 keys x 16 dims per k-step (ldmatrix), B = Q^T registers;
        // two accumulators (even / odd k-steps) halve the dependent HMMA chain
        float sacc[2][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[0][e] = sacc[1][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t af[4];
          const uint32_t off =
              (uint32_t)(((lm & 1) * 8 + lr) * ROWB + (ks * 16 + (lm >> 1) * 8) * (int)sizeof(T));
          ldsm_x4(af, swz128(k_base + off));
          mma_16816<T>(sacc[ks & 1], af, qa[ks][0], qa[ks][1]);
        }
        // lane (g, t): keys {g, g + 8} x rows {2t, 2t + 1}; element 2 hh + h = (key g + 8 hh, row 2t + h)
        float corr[2], pk[2][2];
        bool need_rescale = false;
        const bool plain = !p.use_cap && p.alibi == nullptr;  // no soft cap, no position bias
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float x[2];
          float mx = m[h];
          // interior tile of this row (every key inside its causal / window range): the score is
          // just the scaled dot product — the mask and bias arithmetic is skipped
          const bool inside = plain && pos0 >= row_begin[h] && pos0 + ATT_TILE <= row_end[h];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float sc = sacc[0][2 * hh + h] + sacc[1][2 * hh + h];
            if (inside) {
              x[hh] = sc * p.scale_log2;
            } else {
              const int pos = pos0 + (lane >> 2) + 8 * hh;
              float v = p.use_cap ? tanhf(sc * p.cap_in) * p.cap_out_log2 : sc * p.scale_log2;
              v = fmaf(slope_log2[h], (float)pos, v);
              const bool ok = pos >= row_begin[h] && pos < row_end[h];
              x[hh] = ok ? v : -INFINITY;
            }
            mx = fmaxf(mx, x[hh]);
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));   // over the 8 key groups g
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
          const float ms = (mx == -INFINITY) ? 0.f : mx;
          corr[h] = (mx == m[h]) ? 1.f : exp2f(m[h] - ms);
          pk[0][h] = exp2f(x[0] - ms);
          pk[1][h] = exp2f(x[1] - ms);
          l[h] = fmaf(l[h], corr[h], pk[0][h] + pk[1][h]);
          m[h] = mx;
          need_rescale |= (corr[h] != 1.f);
        }
        if (__any_sync(0xffffffffu, need_rescale)) {
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            o[nb][0] *= corr[0];
            o[nb][1] *= corr[1];
            o[nb][2] *= corr[0];
            o[nb][3] *= corr[1];
          }
        }
        // P^T into the B-operand layout (keys 2t, 2t+1 | +8 of row g): transpose the two 8x8 blocks
        const uint32_t pb0 = movmatrix_trans(Num<T>::pack(pk[0][0], pk[0][1]));
        const uint32_t pb1 = movmatrix_trans(Num<T>::pack(pk[1][0], pk[1][1]));
        // O^T += V^T P^T: A = 16 dims x 16 keys per block (ldmatrix.trans of V's [key][dim] rows)
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
          uint32_t af[4];
          const uint32_t off =
              (uint32_t)(((lm >> 1) * 8 + lr) * ROWB + (mb * 16 + (lm & 1) * 8) * (int)sizeof(T));
          ldsm_x4_trans(af, swz128(v_base + off));
          mma_16816<T>(o[mb], af, pb0, pb1);
        }
        // [tr-emu:tile end]
      } else {
        // [def-emu:tile begin]
        float sacc[2][4];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sacc[nb][e] = 0.f;
#pragma unroll
          for (int kq = 0; kq < KS / 2; ++kq) {
            uint32_t bf[4];
            const uint32_t off = (uint32_t)((nb * 8 + lr) * ROWB + (kq * 32 + lm * 8) * (int)sizeof(T));
            ldsm_x4(bf, swz128(k_base + off));
            mma_16816<T>(sacc[nb], qa[2 * kq], bf[0], bf[1]);
            mma_16816<T>(sacc[nb], qa[2 * kq + 1], bf[2], bf[3]);
          }
        }

        float corr[2];
        bool need_rescale = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float x[4];
          float mx = m[h];
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int pos = pos0 + nb * 8 + (lane & 3) * 2 + e;
              const float sc = sacc[nb][2 * h + e];
              float v = p.use_cap ? tanhf(sc * p.cap_in) * p.cap_out_log2 : sc * p.scale_log2;
              v = fmaf(slope_log2[h], (float)pos, v);
              const bool ok = pos >= row_begin[h] && pos < row_end[h];
              x[nb * 2 + e] = ok ? v : -INFINITY;
              mx = fmaxf(mx, x[nb * 2 + e]);
            }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const float ms = (mx == -INFINITY) ? 0.f : mx;
          corr[h] = (mx == m[h]) ? 1.f : exp2f(m[h] - ms);
          float sum = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            x[c] = exp2f(x[c] - ms);
            sum += x[c];
          }
          l[h] = fmaf(l[h], corr[h], sum);
          m[h] = mx;
          need_rescale |= (corr[h] != 1.f);
          sacc[0][2 * h] = x[0];
          sacc[0][2 * h + 1] = x[1];
          sacc[1][2 * h] = x[2];
          sacc[1][2 * h + 1] = x[3];
        }
        if (__any_sync(0xffffffffu, need_rescale)) {
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            o[nb][0] *= corr[0];
            o[nb][1] *= corr[0];
            o[nb][2] *= corr[1];
            o[nb][3] *= corr[1];
          }
        }
        uint32_t pa[4];
        pa[0] = Num<T>::pack(sacc[0][0], sacc[0][1]);
        pa[1] = Num<T>::pack(sacc[0][2], sacc[0][3]);
        pa[2] = Num<T>::pack(sacc[1][0], sacc[1][1]);
        pa[3] = Num<T>::pack(sacc[1][2], sacc[1][3]);
#pragma unroll
        for (int dq = 0; dq < NB / 2; ++dq) {
          uint32_t bf[4];
          const uint32_t off =
              (uint32_t)(((lm & 1) * 8 + lr) * ROWB + (dq * 16 + (lm >> 1) * 8) * (int)sizeof(T));
          ldsm_x4_trans(bf, swz128(v_base + off));
          mma_16816<T>(o[2 * dq], pa, bf[0], bf[1]);
          mma_16816<T>(o[2 * dq + 1], pa, bf[2], bf[3]);
        }
        // [def-emu:tile end]
      }
      __syncwarp();
      if (boundary) fence_proxy_async_smem();
      ++g_cons;
      issue_ahead();
    }

    // ---- finalize the item: normalised partial O and LSE (log2 domain) ------------------
    if constexpr (TR) {
      // [tr-emu:finalize begin]
      // lane (g, t) owns O^T[dims 16 mb + g, + 8][rows 2t, 2t + 1]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 4);
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 8);
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 16);
        const int r = (lane & 3) * 2 + h;
        if (cur.n_tiles > 0 && r < n_rows) {
          const int row = row0 + r, qi = row / G, head = cur.kvh * G + (row - qi * G);
          const float inv = 1.f / l[h];
          if (p.n_splits == 1) {
            T* dst = static_cast<T*>(p.out) + (int64_t)(cur.q_begin + qi) * p.o_stride_t +
                     (int64_t)head * p.o_stride_h + (lane >> 2);
#pragma unroll
            for (int mb = 0; mb < NB; ++mb) {
              dst[mb * 16] = Num<T>::from_f(o[mb][h] * inv);
              dst[mb * 16 + 8] = Num<T>::from_f(o[mb][2 + h] * inv);
            }
          } else {
            const int64_t wrow = ((int64_t)cur.b * p.max_q_len + qi) * p.n_heads + head;
            float* dst = p.ws_o + (wrow * p.n_splits + cur.split) * D + (lane >> 2);
#pragma unroll
            for (int mb = 0; mb < NB; ++mb) {
              dst[mb * 16] = o[mb][h] * inv;
              dst[mb * 16 + 8] = o[mb][2 + h] * inv;
            }
            if ((lane >> 2) == 0) p.ws_lse[wrow * p.n_splits + cur.split] = m[h] + log2f(l[h]);
          }
        }
      }
      // [tr-emu:finalize end]
    } else {
      // [def-emu:finalize begin]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 1);
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 2);
        const int r = (lane >> 2) + 8 * h;
        if (cur.n_tiles > 0 && r < n_rows) {
          const int row = row0 + r, qi = row / G, head = cur.kvh * G + (row - qi * G);
          const float inv = 1.f / l[h];
          if (p.n_splits == 1) {
            T* dst = static_cast<T*>(p.out) + (int64_t)(cur.q_begin + qi) * p.o_stride_t +
                     (int64_t)head * p.o_stride_h + (lane & 3) * 2;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
              *reinterpret_cast<uint32_t*>(dst + nb * 8) =
                  Num<T>::pack(o[nb][2 * h] * inv, o[nb][2 * h + 1] * inv);
          } else {
            const int64_t wrow = ((int64_t)cur.b * p.max_q_len + qi) * p.n_heads + head;
            float* dst = p.ws_o + (wrow * p.n_splits + cur.split) * D + (lane & 3) * 2;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
              *reinterpret_cast<float2*>(dst + nb * 8) =
                  make_float2(o[nb][2 * h] * inv, o[nb][2 * h + 1] * inv);
            if ((lane & 3) == 0) p.ws_lse[wrow * p.n_splits + cur.split] = m[h] + log2f(l[h]);
          }
        }
      }
      // [def-emu:finalize end]
    }

    // ---- rotate: every pipeline slot advances one stage ---------------------------------------
    cur = nxt;
    cur_issued = nxt_issued;
    nxt_issued = 0;
    if (cur.valid) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int e = 0; e < QR; ++e) qa[ks][e] = qn[ks][e];
      stageC();                       // sB's table window (requested one rotation ago) has landed
      nxt = sB;
      sB = sA;                        // its lengths were requested one rotation ago
      stageB(sB);                     // -> tile range known: request its table window (3rd buffer)
      sA = stageA(s0);
      s0 = stage0();
      if (nxt.valid) load_q(nxt, qn);
      issue_ahead();
    }
  }
}
// [attn-emu:persist end]

// Second pass: merge split-KV partials (same maths as the reference's unwired
// attn_combine_kernel, src/kernels/attention/kernel/attn_combine_kernel.cuh:30).
// One warp per (token, head) row; each lane owns D/32 consecutive outputs (vector loads).
// [attn-emu:combine begin]
template <typename T, int D>
__global__ void __launch_bounds__(128) paged_attn_combine_kernel(const AttnParams p) {
  constexpr int EPL = D / 32;  // elements per lane: 2, 4 or 8 (1 or 3 for head_dim 32 / 96)
  pdl_wait();
  pdl_launch_dependents();  // a following W4A16 GEMM (o_proj) may start prefetching its weights
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x * 4 + warp;
  if (h >= p.n_heads) return;
  const int b = blockIdx.y / p.max_q_len, qi = blockIdx.y % p.max_q_len;
  // the four length words are requested together: one L2 round trip, then LSE, then every partial
  const int q_begin = __ldg(p.q_cu_lens + b), q_end = __ldg(p.q_cu_lens + b + 1);
  const int kv_begin_cu = __ldg(p.kv_cu_lens + b), kv_end_cu = __ldg(p.kv_cu_lens + b + 1);
  const int q_len = q_end - q_begin;
  if (qi >= q_len) return;
  const int64_t tok = q_begin + qi;
  const int64_t row = (int64_t)blockIdx.y * p.n_heads + h;
  const float* lse = p.ws_lse + row * p.n_splits;
  // Lane l looks at split slots l, l+32, ...: is a piece there, and what is its LSE?  Fixed-split
  // kernels publish every slot; the stream kernel only writes pieces that exist, so its static
  // partition is mirrored here.
  auto slot_lse = [&](int s) -> float {
    if (s >= p.n_splits) return -INFINITY;
    if (p.stream) {
      const int G = p.group, kvh = h / G, g = h - kvh * G;
      const int r = qi * G + g, rb = r / 16;
      const int kv_len = kv_end_cu - kv_begin_cu;
      const int rows_total = q_len * G, row0 = rb * 16, n_rows = min(16, rows_total - row0);
      const int q_pos0 = kv_len - q_len, qi_min = row0 / G, qi_max = (row0 + n_rows - 1) / G;
      const int kv_end = q_pos0 + qi_max + 1;
      const int kv_begin = p.window >= 0 ? max(0, q_pos0 + qi_min - p.window) : 0;
      const int64_t seq = ((int64_t)b * p.n_rb + rb) * p.n_kv_heads + kvh;
      const int64_t base = seq * p.ntm, first_piece = base / p.tpw;
      const int64_t lo = max(base, (first_piece + s) * p.tpw);
      const int64_t hi = min(base + p.ntm, (first_piece + s + 1) * p.tpw);
      const int t0 = max((int)(lo - base), kv_begin / ATT_TILE);
      const int t1 = min((int)(hi - base), (kv_end + ATT_TILE - 1) / ATT_TILE);
      if (!(lo < hi && t0 < t1)) return -INFINITY;
    }
    return __ldcg(lse + s);
  };
  const float lse_first = slot_lse(lane);  // kept: the usual case has <= 32 split slots
  float M = lse_first;
  for (int s0 = 32; s0 < p.n_splits; s0 += 32) M = fmaxf(M, slot_lse(s0 + lane));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  float L = 0.f;
  float acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  constexpr int CHUNK = 8;  // partial rows requested per round trip
  for (int s0 = 0; s0 < p.n_splits; s0 += 32) {
    const float my_lse = s0 == 0 ? lse_first : slot_lse(s0 + lane);
    const float my_w = exp2f(my_lse - M);  // 0 for absent / empty pieces
    L += my_w;
    const int ns = min(p.n_splits - s0, 32);
    for (int c0 = 0; c0 < ns; c0 += CHUNK) {
      float w[CHUNK];
      float v[CHUNK][EPL];
#pragma unroll
      for (int c = 0; c < CHUNK; ++c) {
        w[c] = (c0 + c < ns) ? __shfl_sync(0xffffffffu, my_w, (c0 + c) & 31) : 0.f;
        // an absent piece's partial O may be unwritten (NaN): never loaded, never multiplied
        const float* src = p.ws_o + (row * p.n_splits + s0 + c0 + c) * D + lane * EPL;
        if (w[c] != 0.f) {
          if constexpr (EPL == 4) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src));
            v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
          } else if constexpr (EPL == 2) {
            const float2 t = __ldcg(reinterpret_cast<const float2*>(src));
            v[c][0] = t.x; v[c][1] = t.y;
          } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e) v[c][e] = __ldcg(src + e);
          }
        } else {
#pragma unroll
          for (int e = 0; e < EPL; ++e) v[c][e] = 0.f;
        }
      }
#pragma unroll
      for (int c = 0; c < CHUNK; ++c)  // fixed order: deterministic
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = fmaf(v[c][e], w[c], acc[e]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
  const float inv = 1.f / L;
  T* dst = static_cast<T*>(p.out) + tok * p.o_stride_t + (int64_t)h * p.o_stride_h + lane * EPL;
  if constexpr (EPL % 2) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) dst[e] = Num<T>::from_f(acc[e] * inv);
  } else {
#pragma unroll
    for (int e = 0; e < EPL; e += 2)
      *reinterpret_cast<uint32_t*>(dst + e) = Num<T>::pack(acc[e] * inv, acc[e + 1] * inv);
  }
}
// [attn-emu:combine end]

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct MapKey {
  const void* ptr;
  int64_t n_slots, stride_s, stride_h;
  int n_kv_heads, head_dim, box_rows, dtype;
  int mode;  // 0: 3-D, no swizzle (CUDA-core kernel)   1: 4-D {64, D/64, H, slots}, SWIZZLE_128B (mma kernel)
  bool operator==(const MapKey& o) const {
    return mode == o.mode && ptr == o.ptr && n_slots == o.n_slots && stride_s == o.stride_s &&
           stride_h == o.stride_h && n_kv_heads == o.n_kv_heads && head_dim == o.head_dim &&
           box_rows == o.box_rows && dtype == o.dtype;
  }
};
struct MapEntry {
  MapKey key;
  CUtensorMap map;
};
static std::mutex g_map_mu;
static std::vector<MapEntry> g_maps;

static int get_kv_tensor_map(const MapKey& key, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    for (const auto& e : g_maps)
      if (e.key == key) {
        *out = e.map;
        return B200_OK;
      }
  }
  tensor_map_encode_fn enc = get_tensor_map_encode();
  if (!enc) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int es = 2;
  const CUtensorMapDataType dt = key.dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                        : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap m;
  CUresult r;
  if (key.mode == 0) {
    cuuint64_t dims[3] = {(cuuint64_t)key.head_dim, (cuuint64_t)key.n_kv_heads,
                          (cuuint64_t)key.n_slots};
    cuuint64_t strides[2] = {(cuuint64_t)key.stride_h * es, (cuuint64_t)key.stride_s * es};
    cuuint32_t box[3] = {(cuuint32_t)key.head_dim, 1u, (cuuint32_t)key.box_rows};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(&m, dt, 3, const_cast<void*>(key.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {64u, (cuuint64_t)(key.head_dim / 64), (cuuint64_t)key.n_kv_heads,
                          (cuuint64_t)key.n_slots};
    cuuint64_t strides[3] = {128u, (cuuint64_t)key.stride_h * es, (cuuint64_t)key.stride_s * es};
    cuuint32_t box[4] = {64u, (cuuint32_t)(key.head_dim / 64), 1u, (cuuint32_t)key.box_rows};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&m, dt, 4, const_cast<void*>(key.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS)
    return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for kv cache", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.push_back({key, m});
  }
  *out = m;
  return B200_OK;
}

static int ilog2(int x) {
  int s = 0;
  while ((1 << s) < x) ++s;
  return s;
}

// Split-KV plan from host scalars only (CUDA-graph safe).
static void plan_splits(int64_t base_items, int max_kv_len, int ctas_per_sm, int* n_splits,
                        int* tiles_per_split) {
  const int n_tiles = std::max(1, (max_kv_len + ATT_TILE - 1) / ATT_TILE);
  const int64_t slots = (int64_t)sm_count() * ctas_per_sm;
  int s_min = (n_tiles + ATT_MAX_TPS - 1) / ATT_MAX_TPS;
  int s_max = std::max(s_min, std::min(64, n_tiles / 8));  // >= 8 tiles (128 slots) per split
  int best = s_min;
  double best_eff = -1.0;
  const char* env = getenv("B200_ATTN_SPLITS");
  if (env && atoi(env) > 0) {
    best = std::max(s_min, std::min(atoi(env), n_tiles));
  } else {
    for (int s = s_min; s <= s_max; ++s) {
      const double waves = (double)(base_items * s) / (double)slots;
      const double eff = waves / std::ceil(waves);  // tail-wave efficiency
      if (eff > best_eff + 0.02) {                  // prefer fewer splits unless clearly better
        best_eff = eff;
        best = s;
      }
    }
  }
  int tps = (n_tiles + best - 1) / best;
  *tiles_per_split = tps;
  *n_splits = (n_tiles + tps - 1) / tps;
}

static int hg_rows(int group) { return group >= 4 ? 4 : (group >= 2 ? 2 : 1); }

// 0 = CUDA-core kernel ("simt"), 1 = mma.sync kernel, CTA per work item ("mma"),
// 2 = persistent mma.sync kernel (default, "persist").  B200_ATTN_IMPL selects while tuning.
static int attn_impl() {
  const char* e = getenv("B200_ATTN_IMPL");
  if (e && e[0] == 's') return 0;
  if (e && e[0] == 'm') return 1;
  return 2;
}

// B200_ATTN_OCC=1: high-occupancy instantiation of the stream kernel (see ATT_P_TBL_OCC)
static int attn_occ() {
  static const int v = [] {
    const char* e = getenv("B200_ATTN_OCC");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v;
}

// Transposed-tile instantiation of the stream kernel (TR above) for shapes whose packed rows fit
// 8: group * max_q_len <= 8 (decode), head_dim <= 128.  Default since round 2 (91.2 us vs 96.6 us
// at the benchmark shape, profiles/r02_attn_variants.md); B200_ATTN_TR=0 selects the 16-row tile.
static int attn_tr() {
  static const int v = [] {
    const char* e = getenv("B200_ATTN_TR");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

struct AttnPlan {
  int tr;                    // stream kernel: transposed tile, 8-row blocks
  int impl, R, n_hg, n_rb, n_splits, tps, warps;
  int ntm, tpw, n_seq;       // stream kernel
  int64_t total_tiles;
  int64_t grid_y, grid_z;
};

static AttnPlan make_plan(int64_t batch, int max_q_len, int max_kv_len, int n_heads,
                          int n_kv_heads, int head_dim, int block_size) {
  AttnPlan pl{};
  pl.impl = head_dim % 64 ? 0 : attn_impl();  // head_dim 32 / 96: the CUDA-core kernel (no 64-wide swizzle atoms)
  const int group = n_heads / n_kv_heads;
  const int ctas = head_dim <= 128 ? 2 : 1;
  if (pl.impl == 0) {
    pl.R = hg_rows(group);
    pl.n_hg = (group + pl.R - 1) / pl.R;
    pl.n_rb = 1;
    pl.grid_y = (int64_t)n_kv_heads * pl.n_hg;
    pl.grid_z = batch * max_q_len;
  } else {  // both tensor-core kernels pack (q token, group) rows into 16-row blocks
    pl.R = 0;
    pl.n_hg = 1;
    pl.n_rb = (max_q_len * group + 15) / 16;
    pl.grid_y = n_kv_heads;
    pl.grid_z = batch * pl.n_rb;
  }
  pl.warps = ATT_WARPS;
  if (pl.impl == 1) {
    const char* e = getenv("B200_ATTN_WARPS");
    pl.warps = (e && atoi(e) == 4) ? 4 : 1;
  }
  if (pl.impl == 2) {
    pl.warps = 1;
    pl.tr = attn_tr() && head_dim <= 128 && max_q_len * group <= 8;  // => one 8-row block
    pl.ntm = std::max(1, (max_kv_len + ATT_TILE - 1) / ATT_TILE);
    pl.n_seq = (int)(pl.grid_y * pl.grid_z);
    pl.total_tiles = (int64_t)pl.n_seq * pl.ntm;
    const int occ = head_dim <= 128 ? attn_occ() : 0;
    const int64_t warps_resident = (int64_t)sm_count() * (head_dim <= 128 ? (occ ? ATT_OCC_CTAS : 7) : 3);
    int64_t tpw = (pl.total_tiles + warps_resident - 1) / warps_resident;
    const char* e = getenv("B200_ATTN_TPS");
    if (e && atoi(e) > 0) tpw = atoi(e);
    tpw = std::max<int64_t>(tpw, 8);                            // tiny problems: >= 128 slots per piece
    tpw = std::min<int64_t>(tpw, std::max(1, (att_p_tbl(occ) - 8) * block_size / ATT_TILE));  // table window
    tpw = std::min<int64_t>(tpw, 4096);
    pl.tpw = (int)tpw;
    pl.n_splits = (pl.ntm + pl.tpw - 1) / pl.tpw + 1;           // pieces one sequence can be cut into
    if (pl.n_splits > 256) {                                    // absurdly long context with tiny blocks
      pl.impl = 1;
      pl.warps = 4;
      pl.tr = 0;
    } else {
      pl.tps = pl.tpw;
      return pl;
    }
  }
  if (pl.warps == 1) {
    // one warp per CTA: every warp is its own split; ~6 CTAs/SM run at independent phases
    const int n_tiles = std::max(1, (max_kv_len + ATT_TILE - 1) / ATT_TILE);
    const char* e = getenv("B200_ATTN_TPS");
    int tps = 8;
    if (e && atoi(e) > 0) {
      tps = atoi(e);
    } else {
      // largest chunk that still gives every resident warp >= 8 items (dynamic balance within
      // ~1/8 of a warp's work); claims are software pipelined, so small items are cheap
      const int occ = head_dim <= 128 ? attn_occ() : 0;
    const int64_t warps_resident = (int64_t)sm_count() * (head_dim <= 128 ? (occ ? ATT_OCC_CTAS : 7) : 3);
      for (int cand : {32, 16, 8}) {
        tps = cand;
        if (pl.grid_y * pl.grid_z * ((n_tiles + cand - 1) / cand) >= 8 * warps_resident) break;
      }
    }
    // the persistent kernel carries <= 128 block-table entries per item in registers
    if (pl.impl == 2) tps = std::min(tps, std::max(1, 8 * block_size));
    tps = std::max(1, std::min(tps, ATT_TPS_W1));
    // keep the split count (workspace + combine cost) bounded for very long contexts
    while ((n_tiles + tps - 1) / tps > 64 && tps < ATT_TPS_W1) tps *= 2;
    if ((n_tiles + tps - 1) / tps > 256) tps = (n_tiles + 255) / 256;  // > 128K tokens: W=4 path
    if (tps > ATT_TPS_W1) {
      pl.warps = 4;
      if (pl.impl == 2) pl.impl = 1;  // > 128K tokens: CTA-per-item kernel with 4-warp CTAs
    } else {
      pl.tps = tps;
      pl.n_splits = (n_tiles + tps - 1) / tps;
      return pl;
    }
  }
  plan_splits(pl.grid_y * pl.grid_z, max_kv_len, ctas, &pl.n_splits, &pl.tps);
  return pl;
}

// The stream kernel and the combine pass are launched programmatically at B200_PDL level 1 (their
// metadata prologue overlaps the predecessor's tail); B200_ATTN_PDL=0: only at level 2
static int attn_pdl_level() {
  static const int lv = [] {
    const char* e = getenv("B200_ATTN_PDL");
    return (e && e[0] == '0') ? 2 : 1;
  }();
  return lv;
}

template <typename KernelT>
static int launch_kernel(KernelT kernel, size_t smem, int threads, const CUtensorMap& kmap,
                         const CUtensorMap& vmap, const AttnParams& p, const AttnPlan& pl,
                         cudaStream_t st) {
  B200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)p.n_splits, (unsigned)pl.grid_y, (unsigned)pl.grid_z);
  B200_PDL_LAUNCH("paged_attn_decode", kernel, grid, threads, smem, st, kmap, vmap, p);
  return B200_OK;
}

template <typename T, int D>
static int launch_attn(const CUtensorMap& kmap, const CUtensorMap& vmap, const AttnParams& p,
                       const AttnPlan& pl, int64_t batch, cudaStream_t st) {
  constexpr size_t smem = attn_smem_bytes<T, D>();
  int rc;
  if (pl.impl == 2) {
    const unsigned grid = (unsigned)((pl.total_tiles + pl.tpw - 1) / pl.tpw);
    constexpr size_t psmem_occ = (size_t)AttnCfg<D>::STAGES * 2 * ATT_TILE * D * sizeof(T) +
                                 3 * ATT_P_TBL_OCC * sizeof(int32_t) + AttnCfg<D>::STAGES * 8 + 128;
    constexpr size_t psmem_def = (size_t)AttnCfg<D>::STAGES * 2 * ATT_TILE * D * sizeof(T) +
                                 3 * ATT_P_TBL * sizeof(int32_t) + AttnCfg<D>::STAGES * 8 + 128;
    const bool occ = D <= 128 && attn_occ();
    const size_t psmem = occ ? psmem_occ : psmem_def;
    void (*kernel)(const CUtensorMap, const CUtensorMap, const AttnParams, int64_t, int) =
        paged_attn_persist_kernel<T, D, 0>;
    if constexpr (D <= 128) {
      if (occ) kernel = pl.tr ? paged_attn_persist_kernel<T, D, 1, 1> : paged_attn_persist_kernel<T, D, 1>;
      else if (pl.tr) kernel = paged_attn_persist_kernel<T, D, 0, 1>;
    }
    B200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    B200_PDL_LAUNCH_L(attn_pdl_level(), "paged_attn_stream", kernel, grid, 32, psmem, st, kmap, vmap, p,
                      (int64_t)pl.total_tiles, (int)pl.n_seq);
    rc = B200_OK;
  } else if (pl.impl == 1 && pl.warps == 1) {
    rc = launch_kernel(paged_attn_mma_kernel<T, D, 1>, attn_mma_smem_bytes<T, D, 1>(), 32, kmap,
                       vmap, p, pl, st);
  } else if (pl.impl == 1) {
    rc = launch_kernel(paged_attn_mma_kernel<T, D, 4>, attn_mma_smem_bytes<T, D, 4>(), 128, kmap,
                       vmap, p, pl, st);
  } else if (pl.R == 4) {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 4>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  } else if (pl.R == 2) {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 2>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  } else {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 1>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  }
  if (rc != B200_OK) return rc;
  if (p.n_splits > 1) {
    dim3 cgrid((unsigned)((p.n_heads + 3) / 4), (unsigned)(batch * p.max_q_len));
    B200_PDL_LAUNCH_L(attn_pdl_level(), "paged_attn_combine", (paged_attn_combine_kernel<T, D>), cgrid, 128, 0, st, p);
  }
  return B200_OK;
}

// head_dim 32 / 96: only the CUDA-core kernel (and the combine pass) are instantiated
template <typename T, int D>
static int launch_attn_simt(const CUtensorMap& kmap, const CUtensorMap& vmap, const AttnParams& p,
                            const AttnPlan& pl, int64_t batch, cudaStream_t st) {
  constexpr size_t smem = attn_smem_bytes<T, D>();
  int rc;
  if (pl.R == 4) {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 4>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  } else if (pl.R == 2) {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 2>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  } else {
    rc = launch_kernel(paged_attn_decode_kernel<T, D, 1>, smem, ATT_THREADS, kmap, vmap, p, pl, st);
  }
  if (rc != B200_OK) return rc;
  if (p.n_splits > 1) {
    dim3 cgrid((unsigned)((p.n_heads + 3) / 4), (unsigned)(batch * p.max_q_len));
    B200_PDL_LAUNCH_L(attn_pdl_level(), "paged_attn_combine", (paged_attn_combine_kernel<T, D>), cgrid, 128, 0, st, p);
  }
  return B200_OK;
}

template <typename T>
static int launch_attn_d(int D, const CUtensorMap& kmap, const CUtensorMap& vmap,
                         const AttnParams& p, const AttnPlan& pl, int64_t batch, cudaStream_t st) {
  switch (D) {
    case 32: return launch_attn_simt<T, 32>(kmap, vmap, p, pl, batch, st);
    case 96: return launch_attn_simt<T, 96>(kmap, vmap, p, pl, batch, st);
    case 64: return launch_attn<T, 64>(kmap, vmap, p, pl, batch, st);
    case 128: return launch_attn<T, 128>(kmap, vmap, p, pl, batch, st);
    case 256: return launch_attn<T, 256>(kmap, vmap, p, pl, batch, st);
    default:
      return set_error(B200_ERR_UNSUPPORTED, "paged_attn: head_dim %d not in {32,64,96,128,256}", D);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_paged_attn_workspace_bytes(int64_t batch, int64_t max_q_len, int64_t max_kv_len,
                                        int64_t n_heads, int64_t n_kv_heads, int64_t head_dim) {
  if (batch <= 0 || max_q_len <= 0 || n_heads <= 0 || n_kv_heads <= 0) return 0;
  if (n_heads % n_kv_heads) return 0;
  // block_size is not part of this query: size for the block size that needs the most splits
  // (small blocks shrink the block-table window of a piece, but very small ones can also push the
  // plan onto the fixed-split kernel, so no single block size is the worst case)
  int n_splits = 1;
  for (int bs = 1; bs <= 256; bs *= 2) {
    const AttnPlan pl = make_plan(batch, (int)max_q_len, (int)max_kv_len, (int)n_heads,
                                  (int)n_kv_heads, (int)head_dim, bs);
    n_splits = std::max(n_splits, pl.n_splits);
  }
  if (n_splits <= 1) return 0;
  // worst case over env overrides: size for the planned split count
  return batch * max_q_len * n_heads * n_splits * (head_dim + 1) * (int64_t)sizeof(float) + 256;
}

int b200_debug_attn_plan(int64_t batch, int max_q_len, int max_kv_len, int n_heads, int n_kv_heads,
                         int head_dim, int block_size, int64_t* out /*[8]*/) {
  B200_CHECK_ARG(out && batch > 0 && max_q_len > 0 && max_kv_len > 0 && n_heads > 0 && n_kv_heads > 0 &&
                     n_heads % n_kv_heads == 0 && block_size > 0,
                 "debug_attn_plan: bad arguments");
  const AttnPlan pl = make_plan(batch, max_q_len, max_kv_len, n_heads, n_kv_heads, head_dim, block_size);
  out[0] = pl.impl;
  out[1] = pl.n_splits;
  out[2] = pl.tpw;
  out[3] = pl.ntm;
  out[4] = pl.n_seq;
  out[5] = pl.n_rb;
  out[6] = pl.total_tiles;
  out[7] = pl.impl == 2 ? att_p_tbl(head_dim <= 128 ? attn_occ() : 0) : 0;  // table window entries
  return B200_OK;
}

int b200_paged_attn_decode(void* out, const void* q, const void* k_cache, const void* v_cache,
                           const int32_t* q_cu_lens, const int32_t* kv_cu_lens,
                           const int32_t* block_table, const int32_t* block_cu_lens,
                           const float* alibi_slopes, int64_t batch, int64_t n_heads,
                           int64_t n_kv_heads, int64_t head_dim, int64_t n_slots,
                           int64_t q_stride_t, int64_t q_stride_h, int64_t o_stride_t,
                           int64_t o_stride_h, int64_t kv_stride_s, int64_t kv_stride_h,
                           int block_size, int max_q_len, int max_kv_len, float sm_scale,
                           float logits_soft_cap, int sliding_window, void* workspace,
                           int64_t workspace_bytes, int dtype, b200_stream_t stream) {
  B200_CHECK_ARG(out && q && k_cache && v_cache && q_cu_lens && kv_cu_lens && block_table &&
                     block_cu_lens,
                 "paged_attn: null pointer");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_FP16, "paged_attn: dtype must be bf16/fp16");
  B200_CHECK_ARG(batch >= 0 && n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
                 "paged_attn: bad head counts %lld/%lld", (long long)n_heads,
                 (long long)n_kv_heads);
  B200_CHECK_ARG(block_size > 0 && (block_size & (block_size - 1)) == 0,
                 "paged_attn: block_size %d must be a power of two", block_size);
  B200_CHECK_ARG(n_slots > 0 && n_slots < (1ll << 31), "paged_attn: bad n_slots");
  B200_CHECK_ARG(is_aligned(q, 16) && is_aligned(k_cache, 16) && is_aligned(v_cache, 16) &&
                     q_stride_t % 8 == 0 && q_stride_h % 8 == 0 && kv_stride_s % 8 == 0 &&
                     kv_stride_h % 8 == 0,
                 "paged_attn: q / kv cache must be 16-byte aligned with strides %% 8 == 0");
  if (batch == 0 || max_q_len <= 0 || max_kv_len <= 0) return B200_OK;

  const int group = (int)(n_heads / n_kv_heads);
  // prefill / chunked prefill (>= 64 packed query rows per block): the tcgen05 flash kernel of
  // prefill_attn.cu — same operator, no workspace
  if (prefill_attn_eligible(max_q_len, group, (int)head_dim, block_size) && is_aligned(out, 16) &&
      o_stride_t % 8 == 0 && o_stride_h % 8 == 0)
    return launch_prefill_attn(out, q, k_cache, v_cache, q_cu_lens, kv_cu_lens, block_table, block_cu_lens,
                               alibi_slopes, batch, batch * (int64_t)max_q_len, (int)n_heads, (int)n_kv_heads,
                               n_slots, q_stride_t, q_stride_h, o_stride_t, o_stride_h, kv_stride_s,
                               kv_stride_h, block_size, max_q_len, sm_scale, logits_soft_cap, sliding_window,
                               dtype, static_cast<cudaStream_t>(stream));
  const AttnPlan pl = make_plan(batch, max_q_len, max_kv_len, (int)n_heads, (int)n_kv_heads,
                                (int)head_dim, block_size);
  AttnParams p{};
  p.q = q;
  p.out = out;
  p.q_cu_lens = q_cu_lens;
  p.kv_cu_lens = kv_cu_lens;
  p.block_table = block_table;
  p.block_cu_lens = block_cu_lens;
  p.alibi = alibi_slopes;
  p.q_stride_t = q_stride_t;
  p.q_stride_h = q_stride_h;
  p.o_stride_t = o_stride_t;
  p.o_stride_h = o_stride_h;
  p.n_heads = (int)n_heads;
  p.n_kv_heads = (int)n_kv_heads;
  p.group = group;
  p.n_hg = pl.n_hg;
  p.n_rb = pl.n_rb;
  p.block_shift = ilog2(block_size);
  p.block_mask = block_size - 1;
  p.box_rows = block_size < ATT_TILE ? block_size : ATT_TILE;
  p.boxes_per_tile = ATT_TILE / p.box_rows;
  p.max_q_len = max_q_len;
  p.window = sliding_window;
  constexpr float LOG2E = 1.4426950408889634f;
  if (logits_soft_cap > 0.f) {
    p.use_cap = 1;
    p.cap_in = sm_scale / logits_soft_cap;
    p.cap_out_log2 = logits_soft_cap * LOG2E;
  } else {
    p.use_cap = 0;
    p.scale_log2 = sm_scale * LOG2E;
  }
  p.n_splits = pl.n_splits;
  p.tiles_per_split = pl.tps;
  p.ntm = pl.ntm;
  p.tpw = pl.tpw;
  p.stream = pl.impl == 2 ? 1 : 0;
  // workspace layout: [256 B: persistent kernel's work counter][split-KV partial O][partial LSE]
  const int64_t hdr = 0;
  if (p.n_splits > 1) {
    const int64_t rows = batch * max_q_len * n_heads * p.n_splits;
    const int64_t need = hdr + rows * (head_dim + 1) * (int64_t)sizeof(float);
    if (!workspace || workspace_bytes < need)
      return set_error(B200_ERR_WORKSPACE, "paged_attn: workspace %lld B < required %lld B",
                       (long long)workspace_bytes, (long long)need);
    B200_CHECK_ARG(is_aligned(workspace, 16), "paged_attn: workspace must be 16-byte aligned");
    p.ws_o = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + hdr);
    p.ws_lse = p.ws_o + rows * head_dim;
  }

  CUtensorMap kmap, vmap;
  MapKey key{k_cache, n_slots, kv_stride_s, kv_stride_h, (int)n_kv_heads, (int)head_dim,
             p.box_rows, dtype, pl.impl};
  int rc = get_kv_tensor_map(key, &kmap);
  if (rc != B200_OK) return rc;
  key.ptr = v_cache;
  rc = get_kv_tensor_map(key, &vmap);
  if (rc != B200_OK) return rc;

  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_BF16)
    return launch_attn_d<__nv_bfloat16>((int)head_dim, kmap, vmap, p, pl, batch, st);
  return launch_attn_d<__half>((int)head_dim, kmap, vmap, p, pl, batch, st);
}

}  // extern "C"
