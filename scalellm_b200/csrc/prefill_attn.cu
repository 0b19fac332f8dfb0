// prefill_attn.cu — prefill / chunked-prefill paged attention on the 5th-generation tensor cores
// (same operator as the decode kernel: llm::paged_kv_varlen_mha,
// src/kernels/attention/attn_api.cpp:14-73, whose one kernel — Sm80CollectiveMha,
// collective/sm80_collective_mha.cuh:131-402, mma.sync — serves prefill in the reference).
//
// Taken for query blocks of >= 64 packed rows (rows = (query token, head of the kv head's group),
// the reference's own packing of q_len x group into M, sm80_kernel_mha.cuh:208-262), head_dim 128.
// One CTA per (sequence, kv head, block of 128 packed rows); flash attention over 128-key tiles; ten warps:
//   * warp 8 — producer: Q block once (3-D TMA {64 d, G heads, 128/G tokens} x 2, SWIZZLE_128B), then
//     per tile the paged K and V rows as TMA boxes through the block table (first-slot ids,
//     sm80_kernel_mha.cuh:148-152) into two separate 2-stage rings: K(i+2) may land as soon as S(i) has
//     read K(i), V(i+2) once PV(i) has read V(i).  A tile lands as the tcgen05 canonical layouts directly:
//     K = B operand of S = Q K^T, K-major atoms [8 keys x 64 d]; V = B operand of PV, MN-major atoms
//     [8 keys x 64 d] — the same bytes, described differently.  block_size 8: one box of a re-ordered
//     tensor map {64 d, slots, d chunks, heads} carries both d chunks of 8 slots ([chunk][slot][64], atoms
//     2 KB apart); otherwise one box per d chunk of min(block_size, 128) slots;
//   * warp 9 — MMA issuer: S = Q K^T (M 128 rows x N 128 keys x K 128) and PV = P V (M 128 x N 128 d
//     x K 128 keys), tcgen05.mma.kind::f16, fp32 in TMEM; S is double buffered (S0 | S1 | PV: 384
//     columns) and S(i+1) is issued before the wait for P(i), so it runs under the softmax of tile i;
//   * warps 0-7 — softmax, two threads per row (warps w and w + 4 share a TMEM lane quarter; each thread
//     owns 64 keys of the tile and 64 output columns; the row maximum crosses through shared memory once
//     per tile): pass 1 row max, pass 2 exp2 / sum, P rounded to T like the reference
//     (sm80_collective_mha.cuh:289-290) and stored to shared memory as the K-major A operand of the
//     second GEMM; O kept in registers, rescaled by exp2(m_old - m_new) and advanced by each tile's PV
//     read back from TMEM.  Every 32-column chunk is classified per warp: fully visible (one FMNMX per
//     element in pass 1; FFMA + ex2.approx + FADD in pass 2), invisible (skipped) or crossed by the
//     causal diagonal / window edge (masked).  Soft cap / alibi: a separate instantiation (GENERIC).
// Online softmax in the exp2 domain; masks (causal diagonal kv_len - q_len, sliding window, alibi, soft
// cap) exactly as the decode kernel.  No split-KV (a prefill block has enough CTAs), no workspace.
// Per-role cycle traces of the four cuts that led here: profiles/r02_prefill_attn.md (TRACE instantiation,
// tools/prefill_trace.py).
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace b200 {

constexpr int PF_ROWS = 128, PF_KEYS = 128, PF_D = 128;
constexpr int PF_THREADS = 10 * 32;   // 8 softmax warps (two threads per row), producer, MMA issuer
constexpr int PF_WARP_TMA = 8, PF_WARP_MMA = 9;
constexpr int PF_ATOM = 1024;               // 8 rows x 128 B
constexpr int PF_CHUNK = 16 * PF_ATOM;      // 128 rows x 64 elements: 16 KB
constexpr int PF_TILE = 2 * PF_CHUNK;       // 128 x 128 elements: 32 KB
constexpr size_t PF_SMEM = 1024 + (size_t)PF_TILE * (1 /*Q*/ + 2 /*K*/ + 2 /*V*/ + 1 /*P*/) + 16 * 8 + 64 +
                           4 * PF_ROWS * 4 /* row-max exchange */;

struct PrefillParams {
  void* out;
  const int32_t* q_cu_lens;
  const int32_t* kv_cu_lens;
  const int32_t* block_table;
  const int32_t* block_cu_lens;
  const float* alibi;
  int64_t o_stride_t, o_stride_h;
  int n_kv_heads, group, tokens_per_block;  // tokens_per_block = 128 / group
  int block_shift, block_mask, box_rows;
  int dual;                 // 1: one TMA box carries both 64-d chunks of 8 slots ([chunk][slot][64])
  uint32_t kv_sbo;          // bytes between 8-key groups of a K / V tile (1024, dual: 2048)
  uint32_t kv_chunk;        // bytes from a group's d 0..63 atom to its d 64..127 atom (PF_CHUNK, dual: 1024)
  float scale_log2, cap_in, cap_out_log2;
  int use_cap, window;
};

// kind::f16 instruction descriptor: D = f32, A / B of format fmt (0 = f16, 1 = bf16), A K-major,
// B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t pf_idesc(uint32_t M, uint32_t N, uint32_t fmt, uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (0u << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// K-major, 128-byte swizzle, 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t pf_desc_kmajor_sw128(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major, 128-byte swizzle: atoms [8 k x 64 mn]; LBO = bytes between 64-element blocks along MN,
// SBO = bytes between 8-row groups along K (cute::UMMA::make_umma_desc<Major::MN>)
__device__ __forceinline__ uint64_t pf_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float pf_ex2(float x) {   // 2^x, one MUFU; 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, bool TRACE, bool GENERIC>
__global__ void __launch_bounds__(PF_THREADS, 1)
prefill_attn_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap qmap1,
                    const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                    const PrefillParams p, long long* trace) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_smem = base;
  uint8_t* k_smem = q_smem + PF_TILE;        // 2 stages
  uint8_t* v_smem = k_smem + 2 * PF_TILE;    // 2 stages
  uint8_t* p_smem = v_smem + 2 * PF_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + PF_TILE);
  uint64_t* q_full = bars;           // 1
  uint64_t* k_full = bars + 1;       // 2
  uint64_t* v_full = bars + 3;       // 2
  uint64_t* k_empty = bars + 5;      // 2   (S(i) has read K(i))
  uint64_t* v_empty = bars + 7;      // 2   (PV(i) has read V(i))
  uint64_t* s_full = bars + 9;       // 2   (S double buffered in TMEM)
  uint64_t* p_full = bars + 11;      // 1
  uint64_t* pv_full = bars + 12;     // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);
  float* mx_sh = reinterpret_cast<float*>(bars + 16 + 8);   // [parity][half][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  // debug trace (b200_debug_set_trace, tools/prefill_trace.py): clock64 totals per role, 16 slots per CTA
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* tr = (TRACE && trace != nullptr && cta_lin < 1024) ? trace + cta_lin * 16 : nullptr;
  long long tacc[4] = {0, 0, 0, 0};
  long long tt = 0;
  const long long t_start = (TRACE && tr) ? clock64() : 0;
#define PF_T0() do { if (TRACE && tr) tt = clock64(); } while (0)
#define PF_T1(k) do { if (TRACE && tr) { const long long _n = clock64(); tacc[k] += _n - tt; tt = _n; } } while (0)
  const int G = p.group, TPB = p.tokens_per_block;
  // geometry from the step's metadata (inputs of the step: readable before griddepcontrol.wait)
  const int q_begin = p.q_cu_lens[b], q_len = p.q_cu_lens[b + 1] - q_begin;
  const int kv_len = p.kv_cu_lens[b + 1] - p.kv_cu_lens[b];
  const int tok0 = rb * TPB;
  if (tok0 >= q_len) return;  // this sequence has fewer row blocks than the longest one
  const int n_tok = min(TPB, q_len - tok0);
  const int q_pos0 = kv_len - q_len;                       // causal diagonal (sm80_kernel_mha.cuh:258-262)
  const int kv_end = q_pos0 + tok0 + n_tok;                // last row's keys end here
  const int kv_begin = p.window >= 0 ? max(0, q_pos0 + tok0 - p.window) : 0;
  const int t_begin = kv_begin / PF_KEYS, t_end = (kv_end + PF_KEYS - 1) / PF_KEYS;
  const int n_tiles = t_end - t_begin;
  const int blk_cu = p.block_cu_lens[b];

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 2 * PF_ROWS);
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == PF_WARP_MMA) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  if (warp == PF_WARP_TMA && lane == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&qmap1);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t pv_tmem = tmem_base + 256;   // S(i) at tmem_base + (i & 1) * 128
  pdl_launch_dependents();
  constexpr uint32_t FMT = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;  // tcgen05 kind::f16 input format

  if (warp == PF_WARP_TMA) {
    // ===================== producer =====================
    pdl_wait();  // q and the newest KV slots are the predecessor's output
    if (lane == 0) {
      if (n_tok == TPB) {   // one box per 64-d chunk: {64 d, G heads, 128 / G tokens}
        mbar_arrive_expect_tx(q_full, PF_TILE);
        tma_load_3d(q_smem, &qmap, q_full, 0, kvh * G, q_begin + tok0);
        tma_load_3d(q_smem + PF_CHUNK, &qmap, q_full, 64, kvh * G, q_begin + tok0);
      } else {              // ragged last block of a sequence: token by token, never past its tokens
        mbar_arrive_expect_tx(q_full, (uint32_t)(n_tok * G * 128 * 2));
        for (int t = 0; t < n_tok; ++t) {
          tma_load_3d(q_smem + t * G * 128, &qmap1, q_full, 0, kvh * G, q_begin + tok0 + t);
          tma_load_3d(q_smem + PF_CHUNK + t * G * 128, &qmap1, q_full, 64, kvh * G, q_begin + tok0 + t);
        }
      }
    }
    // K(i) may be overwritten once S(i-2) has read it, V(i) once PV(i-2) has: separate rings, so that
    // the K tile of the next S GEMM is on its way while the softmax of the current tile runs.
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      const int pos0 = (t_begin + i) * PF_KEYS;
      const int valid = min(PF_KEYS, kv_end - pos0);                       // keys of this tile that exist
      const int boxes = (valid + p.box_rows - 1) / p.box_rows;             // TMA boxes per chunk per tensor
      const uint32_t bytes = (uint32_t)(boxes * p.box_rows * 128 * 2);     // both 64-d chunks of one tensor
      int slot0[4];                                                        // <= 128 boxes: 4 per lane
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bx = lane + 32 * j;
        const int pos = pos0 + bx * p.box_rows;
        slot0[j] = bx < boxes ? p.block_table[blk_cu + (pos >> p.block_shift)] + (pos & p.block_mask) : 0;
      }
      PF_T0();
      if (lane == 0) {
        mbar_wait(&k_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], bytes);
      }
      __syncwarp();
      PF_T1(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bx = lane + 32 * j;
        if (bx < boxes) {
          uint8_t* kd = k_smem + s * PF_TILE + (bx * p.box_rows >> 3) * p.kv_sbo + ((bx * p.box_rows) & 7) * 128;
          if (p.dual) {   // one box = both chunks of 8 slots: [chunk][slot][64]
            tma_load_4d(kd, &kmap, &k_full[s], 0, slot0[j], 0, kvh);
          } else {
            tma_load_4d(kd, &kmap, &k_full[s], 0, 0, kvh, slot0[j]);
            tma_load_4d(kd + p.kv_chunk, &kmap, &k_full[s], 0, 1, kvh, slot0[j]);
          }
        }
      }
      __syncwarp();
      PF_T1(1);
      if (lane == 0) {
        mbar_wait(&v_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], bytes);
      }
      __syncwarp();
      PF_T1(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bx = lane + 32 * j;
        if (bx < boxes) {
          uint8_t* vd = v_smem + s * PF_TILE + (bx * p.box_rows >> 3) * p.kv_sbo + ((bx * p.box_rows) & 7) * 128;
          if (p.dual) {
            tma_load_4d(vd, &vmap, &v_full[s], 0, slot0[j], 0, kvh);
          } else {
            tma_load_4d(vd, &vmap, &v_full[s], 0, 0, kvh, slot0[j]);
            tma_load_4d(vd + p.kv_chunk, &vmap, &v_full[s], 0, 1, kvh, slot0[j]);
          }
        }
      }
      __syncwarp();
      PF_T1(1);
    }
    if (TRACE && tr && lane == 0) {
      tr[8] = tacc[0];
      tr[9] = tacc[1];
    }
  } else if (warp == PF_WARP_MMA) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = pf_idesc(PF_ROWS, PF_KEYS, FMT, 0);
    constexpr uint32_t idesc_pv = pf_idesc(PF_ROWS, PF_D, FMT, 1);
    const uint32_t q_a = smem_u32(q_smem), k_a = smem_u32(k_smem), v_a = smem_u32(v_smem),
                   p_a = smem_u32(p_smem);
    auto issue_s = [&](int i) {   // S(i) = Q K(i)^T into S buffer i & 1; releases K(i)
      const int s = i & 1;
      mbar_wait(&k_full[s], (i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // over d: 16 per step, 4 steps per 64-d chunk
          const uint32_t sub = (ks & 3) * 32;
          umma_bf16(tmem_base + s * 128, umma_desc_kmajor_sw128(q_a + (ks >> 2) * PF_CHUNK + sub),
                    pf_desc_kmajor_sw128(k_a + s * PF_TILE + (ks >> 2) * p.kv_chunk + sub, p.kv_sbo), idesc_s,
                    ks > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[s]);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    PF_T0();
    mbar_wait(q_full, 0);
    PF_T1(3);
    issue_s(0);
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      PF_T0();
      if (i + 1 < n_tiles) issue_s(i + 1);   // runs on the tensor pipe while the softmax of tile i runs
      PF_T1(0);
      mbar_wait(p_full, i & 1);  // P(i) is in shared memory (and PV(i-1) has been read back)
      mbar_wait(&v_full[s], (i >> 1) & 1);
      PF_T1(1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // PV = P V over keys: 16 per step = two 8-key groups of V
          const uint32_t p_off = (ks >> 2) * PF_CHUNK + (ks & 3) * 32;
          umma_bf16(pv_tmem, umma_desc_kmajor_sw128(p_a + p_off),
                    pf_desc_mnmajor_sw128(v_a + s * PF_TILE + ks * 2 * p.kv_sbo, p.kv_chunk, p.kv_sbo), idesc_pv,
                    ks > 0 ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(&v_empty[s]);
      }
      __syncwarp();
      PF_T1(2);
    }
    if (TRACE && tr && lane == 0) {
      tr[6] = tacc[0];
      tr[7] = tacc[1];
      tr[10] = tacc[3];
      tr[11] = tacc[2];
    }
  } else {
    // ===================== softmax / output: two threads per packed row =====================
    // Thread t: row r = t & 127 (its TMEM lane), column half h = t >> 7 — warps w and w + 4 share a
    // TMEM lane quarter.  Half h owns the tile's keys [64 h, 64 h + 64) (chunk h of P) and the output
    // columns d in [64 h, 64 h + 64).  The two halves agree on the row maximum through shared memory
    // once per tile; their partial row sums are added at the end.
    const int r = threadIdx.x & 127;
    const int h = threadIdx.x >> 7;
    const int qi = r / G, g = r - qi * G;          // token inside the block, head inside the group
    const bool row_ok = qi < n_tok;
    const int head = kvh * G + g;
    const int row_end = row_ok ? q_pos0 + tok0 + qi + 1 : 0;               // keys [row_begin, row_end)
    const int row_begin = p.window >= 0 ? max(0, q_pos0 + tok0 + qi - p.window) : 0;
    const float slope_log2 = p.alibi ? p.alibi[head] * 1.4426950408889634f : 0.f;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int pair_bar = 1 + (warp & 3);           // named barrier of warps w and w + 4 (64 threads)
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, corr = 1.f;
    auto score = [&](uint32_t raw, int pos) -> float {
      const float acc = __uint_as_float(raw);
      float v = p.use_cap ? tanhf(acc * p.cap_in) * p.cap_out_log2 : acc * p.scale_log2;
      v = fmaf(slope_log2, (float)pos, v);
      return (pos >= row_begin && pos < row_end) ? v : -INFINITY;
    };
    // 32 keys = four 16-byte units of row r inside chunk h of P (K-major, 128-byte swizzle); c = 0 / 1
    auto store_p = [&](int c, const uint32_t (&pk)[16]) {
      uint8_t* prow = p_smem + h * PF_CHUNK + r * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int unit = c * 4 + u;
        *reinterpret_cast<uint4*>(prow + ((unit ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
    };
    constexpr bool generic = GENERIC;   // soft cap or alibi: the per-element score path (host-selected)
    auto add_pv = [&](float c) {   // o = o * c + PV (the tile whose PV is in TMEM), this half's columns
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t y[32];
        tmem_ld_32x32b_x32(pv_tmem + lane_addr + h * 64 + c0, y);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c0 + i] = fmaf(o[c0 + i], c, __uint_as_float(y[i]));
      }
    };
    for (int i = 0; i < n_tiles; ++i) {
      const int pos0 = (t_begin + i) * PF_KEYS;
      const uint32_t s_tmem = tmem_base + (i & 1) * 128 + lane_addr + h * 64;   // this thread's 64 columns
      const int col0 = h * 64;
      PF_T0();
      mbar_wait(&s_full[i & 1], (i >> 1) & 1);
      PF_T1(0);
      tc_fence_after();
      // Columns of this tile that row r may see: [lo, hi) (causal diagonal, sliding window).  Without
      // soft cap / alibi a 32-column chunk that is fully visible to every row of the warp (all chunks of
      // an interior tile) costs one FMNMX per element in pass 1 and FFMA + ex2 + add in pass 2; a chunk
      // no row of the warp sees is skipped; only chunks the diagonal crosses pay for the masks.
      const int lo = row_begin - pos0, hi = row_end - pos0;
      const unsigned span = hi > lo ? (unsigned)(hi - lo) : 0u;
      uint32_t full_mask = 0, none_mask = 0;   // bit c: this half's chunk c fully visible / invisible (warp-uniform)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cb = col0 + c * 32;
        if (__all_sync(0xffffffffu, lo <= cb && hi >= cb + 32)) full_mask |= 1u << c;
        if (__all_sync(0xffffffffu, hi <= cb || lo >= cb + 32)) none_mask |= 1u << c;
      }
      // ---- pass 1: row maximum ----
      float mx_part = -INFINITY;
      if (none_mask != 3u) {
        uint32_t x[64];   // this thread's 64 scores (raw accumulators)
        tmem_ld_32x32b_x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&x[0]));
        tmem_ld_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&x[32]));
        tmem_ld_wait();
        if (generic) {
#pragma unroll
          for (int j = 0; j < 64; ++j) mx_part = fmaxf(mx_part, score(x[j], pos0 + col0 + j));
        } else {
          float r0 = -INFINITY, r1 = -INFINITY, r2 = -INFINITY, r3 = -INFINITY;   // raw accumulator maxima
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if ((full_mask >> c) & 1u) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                r0 = fmaxf(r0, __uint_as_float(x[c * 32 + j]));
                r1 = fmaxf(r1, __uint_as_float(x[c * 32 + j + 1]));
                r2 = fmaxf(r2, __uint_as_float(x[c * 32 + j + 2]));
                r3 = fmaxf(r3, __uint_as_float(x[c * 32 + j + 3]));
              }
            } else if (!((none_mask >> c) & 1u)) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if ((unsigned)(col0 + c * 32 + j - lo) < span) r0 = fmaxf(r0, __uint_as_float(x[c * 32 + j]));
            }
          }
          mx_part = fmaxf(fmaxf(r0, r1), fmaxf(r2, r3)) * p.scale_log2;   // scale > 0: max commutes
        }
      }
      // the other half's maximum (parity-buffered: a thread cannot be two tiles ahead of its partner)
      mx_sh[((i & 1) * 2 + h) * PF_ROWS + r] = mx_part;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float mx = fmaxf(m, fmaxf(mx_part, mx_sh[((i & 1) * 2 + (h ^ 1)) * PF_ROWS + r]));
      PF_T1(1);
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      const float corr_new = exp2f(m - ms);         // rescales everything accumulated so far
      // the previous tile's PV must be folded in (and P's buffer released by its MMAs) before P is rewritten
      if (i > 0) {
        mbar_wait(pv_full, (i - 1) & 1);
        tc_fence_after();
        add_pv(corr);
      }
      corr = corr_new;
      PF_T1(2);
      // ---- pass 2: p = exp2(s - m), row sum, P -> T -> shared memory ----
      float s0 = 0.f, s1 = 0.f;
      const float neg_ms = -ms;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
        uint32_t x[32];   // loaded again rather than kept across the fold above (registers)
        if (!((none_mask >> c) & 1u)) {
          tmem_ld_32x32b_x32(s_tmem + c * 32, x);
          tmem_ld_wait();
        }
        if ((none_mask >> c) & 1u) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        } else if (generic) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = exp2f(score(x[j], pos0 + col0 + c * 32 + j) - ms);
            const float p1 = exp2f(score(x[j + 1], pos0 + col0 + c * 32 + j + 1) - ms);
            pk[j >> 1] = Num<T>::pack(p0, p1);   // P rounded to T for the second GEMM ...
            s0 += p0;                             // ... the row sum stays fp32 (online_softmax.cuh:39-162)
            s1 += p1;
          }
        } else if ((full_mask >> c) & 1u) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = pf_ex2(fmaf(__uint_as_float(x[j]), p.scale_log2, neg_ms));
            const float p1 = pf_ex2(fmaf(__uint_as_float(x[j + 1]), p.scale_log2, neg_ms));
            pk[j >> 1] = Num<T>::pack(p0, p1);
            s0 += p0;
            s1 += p1;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0 = pf_ex2(fmaf(__uint_as_float(x[j]), p.scale_log2, neg_ms));
            float p1 = pf_ex2(fmaf(__uint_as_float(x[j + 1]), p.scale_log2, neg_ms));
            p0 = (unsigned)(col0 + c * 32 + j - lo) < span ? p0 : 0.f;
            p1 = (unsigned)(col0 + c * 32 + j + 1 - lo) < span ? p1 : 0.f;
            pk[j >> 1] = Num<T>::pack(p0, p1);
            s0 += p0;
            s1 += p1;
          }
        }
        store_p(c, pk);
      }
      l = fmaf(l, corr_new, s0 + s1);   // this half's share of the row sum
      m = mx;
      // keys of the tile that do not exist: their V rows hold whatever the ring held (maybe NaN) and
      // 0 * NaN is NaN — zero them (their P columns are exactly 0 already); half h clears d chunk h
      const int valid = min(PF_KEYS, kv_end - pos0);
      if (valid < PF_KEYS) {   // CTA-uniform: the sequence's last tile
        mbar_wait(&v_full[i & 1], (i >> 1) & 1);   // the boxes that do exist have landed
        if (r >= valid) {
          uint8_t* vrow = v_smem + (i & 1) * PF_TILE + (r >> 3) * p.kv_sbo + (r & 7) * 128 + h * p.kv_chunk;
#pragma unroll
          for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(vrow + u * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      PF_T1(3);
    }
    if (TRACE && tr && threadIdx.x == 0) {
      tr[1] = n_tiles;
      tr[2] = tacc[0];
      tr[3] = tacc[1];
      tr[4] = tacc[2];
      tr[5] = tacc[3];
    }
    mbar_wait(pv_full, (n_tiles - 1) & 1);
    tc_fence_after();
    add_pv(corr);
    // the row sum: both halves' shares
    mx_sh[h * PF_ROWS + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    l += mx_sh[(h ^ 1) * PF_ROWS + r];
    if (row_ok) {
      const float inv = 1.f / l;
      T* dst = static_cast<T*>(p.out) + (int64_t)(q_begin + tok0 + qi) * p.o_stride_t + (int64_t)head * p.o_stride_h + h * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        uint4 v;
        v.x = Num<T>::pack(o[c0] * inv, o[c0 + 1] * inv);
        v.y = Num<T>::pack(o[c0 + 2] * inv, o[c0 + 3] * inv);
        v.z = Num<T>::pack(o[c0 + 4] * inv, o[c0 + 5] * inv);
        v.w = Num<T>::pack(o[c0 + 6] * inv, o[c0 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c0) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TRACE && tr && threadIdx.x == 0) tr[0] = clock64() - t_start;
  if (warp == PF_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#undef PF_T0
#undef PF_T1
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct PfMapKey {
  const void* ptr;
  int64_t d0, d1, d2, s1, s2;   // dims / strides (elements) of the 3-D or 4-D view
  int box1, box2, dtype, kind;  // kind 0: q {D, H, T} box {64, box1, box2}; 1: kv {64, D/64, H, slots} box {64,1,1,box2}
  bool operator==(const PfMapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 &&
           box1 == o.box1 && box2 == o.box2 && dtype == o.dtype && kind == o.kind;
  }
};
static std::mutex g_pf_mu;
static std::vector<std::pair<PfMapKey, CUtensorMap>> g_pf_maps;

static int pf_get_map(const PfMapKey& k, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_pf_mu);
    for (const auto& e : g_pf_maps)
      if (e.first == k) {
        *out = e.second;
        return B200_OK;
      }
  }
  tensor_map_encode_fn enc = get_tensor_map_encode();
  if (!enc) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const CUtensorMapDataType dt = k.dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap m;
  CUresult r;
  if (k.kind == 0) {   // q [T, H, D]: dims {D, H, T}, strides {s1 = head stride, s2 = token stride}
    cuuint64_t dims[3] = {(cuuint64_t)k.d0, (cuuint64_t)k.d1, (cuuint64_t)k.d2};
    cuuint64_t strides[2] = {(cuuint64_t)k.s1 * 2, (cuuint64_t)k.s2 * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)k.box1, (cuuint32_t)k.box2};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(&m, dt, 3, const_cast<void*>(k.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (k.kind == 2) {   // kv cache as {64, slots, D/64, Hkv}: one box = {64, 8 slots, both chunks, 1 head},
                              // landing as [chunk][slot][64] — half the TMA operations at block_size 8
    cuuint64_t dims[4] = {64u, (cuuint64_t)k.d2, (cuuint64_t)(k.d0 / 64), (cuuint64_t)k.d1};
    cuuint64_t strides[3] = {(cuuint64_t)k.s2 * 2, 128u, (cuuint64_t)k.s1 * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)k.box2, (cuuint32_t)(k.d0 / 64), 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&m, dt, 4, const_cast<void*>(k.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {             // kv cache [slots, Hkv, D] as {64, D/64, Hkv, slots}: one 64-wide chunk per box
    cuuint64_t dims[4] = {64u, (cuuint64_t)(k.d0 / 64), (cuuint64_t)k.d1, (cuuint64_t)k.d2};
    cuuint64_t strides[3] = {128u, (cuuint64_t)k.s1 * 2, (cuuint64_t)k.s2 * 2};
    cuuint32_t box[4] = {64u, 1u, 1u, (cuuint32_t)k.box2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&m, dt, 4, const_cast<void*>(k.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for prefill attention", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_pf_mu);
    if (g_pf_maps.size() > 4096) g_pf_maps.clear();
    g_pf_maps.push_back({k, m});
  }
  *out = m;
  return B200_OK;
}

bool prefill_attn_eligible(int max_q_len, int group, int head_dim, int block_size) {
  static const bool off = [] {
    const char* e = getenv("B200_ATTN_PREFILL");
    return e && e[0] == '0';
  }();
  return !off && head_dim == PF_D && group >= 1 && group <= PF_ROWS && (PF_ROWS % group) == 0 &&
         (int64_t)max_q_len * group >= 64 && block_size >= 1;
}

int launch_prefill_attn(void* out, const void* q, const void* k_cache, const void* v_cache,
                        const int32_t* q_cu_lens, const int32_t* kv_cu_lens, const int32_t* block_table,
                        const int32_t* block_cu_lens, const float* alibi, int64_t batch, int64_t n_tokens_bound,
                        int n_heads, int n_kv_heads, int64_t n_slots, int64_t q_stride_t, int64_t q_stride_h,
                        int64_t o_stride_t, int64_t o_stride_h, int64_t kv_stride_s, int64_t kv_stride_h,
                        int block_size, int max_q_len, float sm_scale, float soft_cap, int window, int dtype,
                        cudaStream_t st) {
  const int G = n_heads / n_kv_heads;
  PrefillParams p{};
  p.out = out;
  p.q_cu_lens = q_cu_lens;
  p.kv_cu_lens = kv_cu_lens;
  p.block_table = block_table;
  p.block_cu_lens = block_cu_lens;
  p.alibi = alibi;
  p.o_stride_t = o_stride_t;
  p.o_stride_h = o_stride_h;
  p.n_kv_heads = n_kv_heads;
  p.group = G;
  p.tokens_per_block = PF_ROWS / G;
  int sh = 0;
  while ((1 << sh) < block_size) ++sh;
  p.block_shift = sh;
  p.block_mask = block_size - 1;
  // K / V boxes: block_size 8 (the serving default) -> one box per 8 slots carrying both d chunks, if the
  // driver takes the re-ordered tensor map (B200_ATTN_PF_DUAL=0 disables); otherwise one box per chunk
  // of min(block_size, 128) slots
  static const bool dual_ok = [] {
    const char* e = getenv("B200_ATTN_PF_DUAL");
    return !(e && e[0] == '0');
  }();
  p.dual = (block_size == 8 && dual_ok) ? 1 : 0;
  p.box_rows = p.dual ? 8 : (block_size < PF_KEYS ? block_size : PF_KEYS);
  constexpr float LOG2E = 1.4426950408889634f;
  if (soft_cap > 0.f) {
    p.use_cap = 1;
    p.cap_in = sm_scale / soft_cap;
    p.cap_out_log2 = soft_cap * LOG2E;
  } else {
    p.scale_log2 = sm_scale * LOG2E;
  }
  p.window = window;
  CUtensorMap qmap, qmap1, kmap, vmap;
  // q: the extent is an upper bound of the token count (batch * max_q_len); no box ever reaches past
  // a sequence's own tokens (full blocks use the block-sized box, ragged ones the one-token box)
  int rc = pf_get_map(PfMapKey{q, PF_D, n_heads, n_tokens_bound, q_stride_h, q_stride_t, G, p.tokens_per_block, dtype, 0}, &qmap);
  if (rc != B200_OK) return rc;
  rc = pf_get_map(PfMapKey{q, PF_D, n_heads, n_tokens_bound, q_stride_h, q_stride_t, G, 1, dtype, 0}, &qmap1);
  if (rc != B200_OK) return rc;
  if (p.dual) {
    rc = pf_get_map(PfMapKey{k_cache, PF_D, n_kv_heads, n_slots, kv_stride_h, kv_stride_s, 1, 8, dtype, 2}, &kmap);
    if (rc == B200_OK)
      rc = pf_get_map(PfMapKey{v_cache, PF_D, n_kv_heads, n_slots, kv_stride_h, kv_stride_s, 1, 8, dtype, 2}, &vmap);
    if (rc != B200_OK) {   // the driver refused the re-ordered map: per-chunk boxes
      p.dual = 0;
      p.box_rows = block_size < PF_KEYS ? block_size : PF_KEYS;
    }
  }
  if (!p.dual) {
    rc = pf_get_map(PfMapKey{k_cache, PF_D, n_kv_heads, n_slots, kv_stride_h, kv_stride_s, 1, p.box_rows, dtype, 1}, &kmap);
    if (rc != B200_OK) return rc;
    rc = pf_get_map(PfMapKey{v_cache, PF_D, n_kv_heads, n_slots, kv_stride_h, kv_stride_s, 1, p.box_rows, dtype, 1}, &vmap);
    if (rc != B200_OK) return rc;
  }
  p.kv_sbo = p.dual ? 2048u : 1024u;
  p.kv_chunk = p.dual ? 1024u : (uint32_t)PF_CHUNK;
  const unsigned n_rb = (unsigned)(((int64_t)max_q_len + p.tokens_per_block - 1) / p.tokens_per_block);
  dim3 grid(n_rb, (unsigned)n_kv_heads, (unsigned)batch);
  if (dtype == B200_BF16) {
    const bool gen = p.use_cap != 0 || p.alibi != nullptr;
    auto kern = debug_trace_ptr() ? (gen ? prefill_attn_kernel<__nv_bfloat16, true, true> : prefill_attn_kernel<__nv_bfloat16, true, false>)
                                  : (gen ? prefill_attn_kernel<__nv_bfloat16, false, true> : prefill_attn_kernel<__nv_bfloat16, false, false>);
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM));
    B200_PDL_LAUNCH_L(1, "prefill_attn", kern, grid, PF_THREADS, PF_SMEM, st, qmap, qmap1, kmap, vmap, p, debug_trace_ptr());
  } else {
    const bool gen = p.use_cap != 0 || p.alibi != nullptr;
    auto kern = debug_trace_ptr() ? (gen ? prefill_attn_kernel<__half, true, true> : prefill_attn_kernel<__half, true, false>)
                                  : (gen ? prefill_attn_kernel<__half, false, true> : prefill_attn_kernel<__half, false, false>);
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM));
    B200_PDL_LAUNCH_L(1, "prefill_attn", kern, grid, PF_THREADS, PF_SMEM, st, qmap, qmap1, kmap, vmap, p, debug_trace_ptr());
  }
  return B200_OK;
}

}  // namespace b200
