// attn_emu.cpp — host execution of the WHOLE paged-attention stream kernel
// (paged_attn_persist_kernel in scalellm_b200/csrc/paged_attn.cu) on a paged KV cache.
//
// tools/attn_emu.py cuts the kernel and the definitions it needs out of the .cu file
// ([attn-emu:params], [attn-emu:persist]), rewrites the two inline cp.async statements into calls
// of this harness, and compiles it once per instantiation (-DEMU_OCC, -DEMU_TR).  Every CTA of the
// grid is run by 32 host threads (the lanes of its one warp) executing the kernel's own statements:
// the metadata pipeline, the TMA ring, the tile loop and the output / partial scatter.  Emulated:
// mbarriers, the 4-D tensor-map loads into 128-byte-swizzled shared memory (out-of-range slots
// read as zero), cp.async, ldmatrix, mma.sync, movmatrix, shuffles and votes.  The work partition
// comes from the library itself (b200_debug_attn_plan).  The partials are merged with the LSE
// formula and compared with a plain softmax(QK^T)V per sequence; with split KV the combine kernel
// (cut out of the source as well, [attn-emu:combine]) runs after the stream kernel and the final
// bf16 output is compared too.  The default instantiation is
// validated on the GPU, so it validates this harness; the opt-in ones (OCC, TR) are then checked
// by the same harness.
#include <pthread.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#ifndef EMU_OCC
#define EMU_OCC 0
#endif
#ifndef EMU_TR
#define EMU_TR 0
#endif
#ifndef EMU_D
#define EMU_D 128
#endif
#ifndef EMU_SIMT  // 1: the CUDA-core kernel (paged_attn_decode_kernel, 4 warps per CTA) instead of the stream kernel
#define EMU_SIMT 0
#endif

extern "C" int b200_debug_attn_plan(int64_t batch, int max_q_len, int max_kv_len, int n_heads, int n_kv_heads,
                                    int head_dim, int block_size, int64_t* out);

// ---- CUDA-isms -------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__
#define __align__(n)
using std::max;
using std::min;
struct Dim3 {
  unsigned x, y, z;
};
static thread_local Dim3 blockIdx, threadIdx;
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }

struct bf16_t {
  uint16_t bits;
};
static inline float bf2f(bf16_t x) {
  uint32_t u = (uint32_t)x.bits << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline bf16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return bf16_t{(uint16_t)((u >> 16) | 0x40)};
  u += 0x7fffu + ((u >> 16) & 1u);
  return bf16_t{(uint16_t)(u >> 16)};
}
template <typename T>
struct Num;
template <>
struct Num<bf16_t> {
  static float to_f(bf16_t x) { return bf2f(x); }
  static bf16_t from_f(float x) { return f2bf(x); }
  static uint32_t pack(float lo, float hi) { return (uint32_t)f2bf(lo).bits | ((uint32_t)f2bf(hi).bits << 16); }
  static float2 unpack(uint32_t w) { return float2{bf2f(bf16_t{(uint16_t)(w & 0xffffu)}), bf2f(bf16_t{(uint16_t)(w >> 16)})}; }
};

// the KV cache as the tensor map describes it: [n_slots][n_kv_heads][D]
struct CUtensorMap {
  const bf16_t* base;
  int64_t n_slots;
  int n_kv_heads, head_dim, box_rows;
};

// ---- one warp = 32 threads -----------------------------------------------------------------------------
constexpr int LANES = 32;
struct WarpCtx {  // what the lanes of one warp share to emulate a warp-collective instruction
  pthread_barrier_t bar;
  uint32_t x32[LANES][8];
  float xf[LANES];
  int xi[LANES];
};
static WarpCtx g_warps[4];
static thread_local WarpCtx* t_warp = &g_warps[0];
static pthread_barrier_t g_cta_bar;  // __syncthreads of a multi-warp CTA
static thread_local int t_lane;
uint8_t smem_raw[160 * 1024] __attribute__((aligned(1024)));
static std::mutex g_mu;
static bool g_failed = false;
static void fail(const char* what) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_failed) std::fprintf(stderr, "attn_emu: %s\n", what);
  g_failed = true;
}
static void sync_warp() { pthread_barrier_wait(&t_warp->bar); }
static inline void __syncthreads() { pthread_barrier_wait(&g_cta_bar); }
static inline void __syncwarp() { sync_warp(); }
static inline void pdl_wait() {}
static inline void pdl_launch_dependents() {}
static inline void fence_mbar_init() {}
static inline void fence_proxy_async_smem() {}
static inline void prefetch_tensormap(const CUtensorMap*) {}
static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - smem_raw); }

static inline float __shfl_xor_sync(unsigned, float v, int mask) {
  t_warp->xf[t_lane] = v;
  sync_warp();
  const float r = t_warp->xf[t_lane ^ mask];
  sync_warp();
  return r;
}
static inline bool __any_sync(unsigned, bool pred) {
  t_warp->xi[t_lane] = pred;
  sync_warp();
  bool r = false;
  for (int i = 0; i < LANES; ++i) r |= t_warp->xi[i] != 0;
  sync_warp();
  return r;
}
static inline uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }

static void ldsm_impl(uint32_t (&r)[4], uint32_t addr, bool trans) {
  t_warp->x32[t_lane][0] = addr;
  sync_warp();
  for (int j = 0; j < 4; ++j) {
    uint16_t e[2];
    for (int k = 0; k < 2; ++k) {
      const int row = trans ? 2 * (t_lane & 3) + k : t_lane >> 2;
      const int col = trans ? t_lane >> 2 : 2 * (t_lane & 3) + k;
      std::memcpy(&e[k], smem_raw + t_warp->x32[8 * j + row][0] + col * 2, 2);
    }
    r[j] = (uint32_t)e[0] | ((uint32_t)e[1] << 16);
  }
  sync_warp();
}
static void ldsm_x4(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, false); }
static void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, true); }
static uint32_t movmatrix_trans(uint32_t a) {
  t_warp->x32[t_lane][0] = a;
  sync_warp();
  uint16_t e[2];
  for (int k = 0; k < 2; ++k) {
    const int src_row = 2 * (t_lane & 3) + k, src_col = t_lane >> 2;
    const uint32_t w = t_warp->x32[src_row * 4 + src_col / 2][0];
    e[k] = (uint16_t)(src_col & 1 ? w >> 16 : w & 0xffffu);
  }
  sync_warp();
  return (uint32_t)e[0] | ((uint32_t)e[1] << 16);
}
template <typename T>
static void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  for (int i = 0; i < 4; ++i) t_warp->x32[t_lane][i] = a[i];
  t_warp->x32[t_lane][4] = b0;
  t_warp->x32[t_lane][5] = b1;
  sync_warp();
  auto half = [](uint32_t w, int k) { return bf2f(bf16_t{(uint16_t)(k ? w >> 16 : w & 0xffffu)}); };
  auto A = [&](int row, int col) {
    const int g = row & 7, t = (col & 7) >> 1;
    return half(t_warp->x32[g * 4 + t][(row >> 3) + 2 * (col >> 3)], col & 1);
  };
  auto B = [&](int k, int n) { return half(t_warp->x32[n * 4 + ((k & 7) >> 1)][4 + (k >> 3)], k & 1); };
  const int g = t_lane >> 2, t = t_lane & 3;
  const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  float acc[4];
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    for (int k = 0; k < 16; ++k) s += (double)A(rows[i], k) * (double)B(k, cols[i]);
    acc[i] = d[i] + (float)s;
  }
  sync_warp();
  for (int i = 0; i < 4; ++i) d[i] = acc[i];
}

// ---- mbarriers + TMA + cp.async --------------------------------------------------------------------------
struct EmuBar {
  int count = 0, pending = 0, phase = 0;
  long tx = 0;
};
static EmuBar g_eb[16];
static uint64_t* g_bar_base = nullptr;  // the kernel's barrier array inside smem_raw
static EmuBar& eb(uint64_t* bar) {
  if (!g_bar_base) g_bar_base = bar;
  return g_eb[bar - g_bar_base];
}
static void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_bar_base || bar < g_bar_base) g_bar_base = bar - threadIdx.x;  // thread i initialises bars[i]
  EmuBar& b = g_eb[bar - g_bar_base];
  b.count = b.pending = (int)count;
  b.phase = 0;
  b.tx = 0;
}
static void bar_check(EmuBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    ++b.phase;
    b.pending = b.count;
  }
}
static void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  EmuBar& b = eb(bar);
  b.tx += bytes;
  --b.pending;
  bar_check(b);
}
static void mbar_wait(uint64_t* bar, uint32_t parity) {
  // wall-clock deadline (a loaded machine may not schedule lane 0 for a long while), with back-off
  const auto t0 = std::chrono::steady_clock::now();
  for (int spin = 0;; ++spin) {
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if ((uint32_t)(eb(bar).phase & 1) != parity) return;
      if (g_failed) return;
    }
    if (spin < 200) {
      sched_yield();
    } else {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      if ((spin & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) {
        fail("mbar_wait never satisfied (the ring stalled)");
        return;
      }
    }
  }
}
// box {64, D/64, 1, box_rows} at (0, 0, kvh, slot0): rows slot0.. of head kvh, each D elements, land
// row-major [row][D] at dst with the 128-byte swizzle; slots past the end read as zero
static void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int, int, int kvh, int slot0) {
  const int D = map->head_dim;
  const uint32_t base = smem_u32(dst);
  for (int r = 0; r < map->box_rows; ++r)
    for (int d = 0; d < D; ++d) {
      bf16_t v = f2bf(0.f);
      const int64_t slot = (int64_t)slot0 + r;
      if (slot >= 0 && slot < map->n_slots) v = map->base[(slot * map->n_kv_heads + kvh) * D + d];
      std::memcpy(smem_raw + swz128(base + (uint32_t)(r * D + d) * 2), &v, 2);
    }
  std::lock_guard<std::mutex> lk(g_mu);
  EmuBar& b = eb(bar);
  b.tx -= (long)map->box_rows * D * 2;
  bar_check(b);
}
static void emu_cp_async4(uint32_t dst, const void* src) { std::memcpy(smem_raw + dst, src, 4); }
static void emu_cp_async_wait_all() {}

struct float4 {
  float x, y, z, w;
};
template <typename V>
static inline V __ldg(const V* p) { return *p; }
template <typename V>
static inline V __ldcg(const V* p) { return *p; }
static inline float __shfl_sync(unsigned, float v, int src) {
  t_warp->xf[t_lane] = v;
  sync_warp();
  const float r = t_warp->xf[src & 31];
  sync_warp();
  return r;
}

static inline uint4 ld_v4(const void* p) {
  uint4 r;
  std::memcpy(&r, p, 16);
  return r;
}
// box {D, 1, box_rows} at (0, kvh, slot0), no swizzle: rows land contiguously [row][D]
static void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int, int kvh, int slot0) {
  const int D = map->head_dim;
  for (int r = 0; r < map->box_rows; ++r)
    for (int d = 0; d < D; ++d) {
      bf16_t v = f2bf(0.f);
      const int64_t slot = (int64_t)slot0 + r;
      if (slot >= 0 && slot < map->n_slots) v = map->base[(slot * map->n_kv_heads + kvh) * D + d];
      std::memcpy((uint8_t*)dst + (size_t)(r * D + d) * 2, &v, 2);
    }
  std::lock_guard<std::mutex> lk(g_mu);
  EmuBar& b = eb(bar);
  b.tx -= (long)map->box_rows * D * 2;
  bar_check(b);
}

#include "attn_emu_params.inc"
#include "attn_emu_simt.inc"
#include "attn_emu_persist.inc"
#include "attn_emu_combine.inc"

// ---- one problem ---------------------------------------------------------------------------------------------
struct Problem {
  int B, H, Hkv, D, bs, max_q;
  std::vector<int> q_lens, kv_lens;
  std::vector<int32_t> q_cu, kv_cu, blk_cu, table;
  std::vector<bf16_t> q, out, kc, vc;
  std::vector<float> ws_o, ws_lse;
  int64_t n_slots;
  AttnParams p;
  CUtensorMap kmap, vmap;
  int64_t total_tiles;
  int n_seq;
};
static Problem* g_pr;

static void* lane_main(void* arg) {
  t_lane = (int)(intptr_t)arg;
  threadIdx.x = (unsigned)t_lane;
#if !EMU_SIMT
  Problem& P = *g_pr;
  paged_attn_persist_kernel<bf16_t, EMU_D, EMU_OCC, EMU_TR>(P.kmap, P.vmap, P.p, P.total_tiles, P.n_seq);
#endif
  return nullptr;
}

static int run(int B, int H, int Hkv, int bs, int max_q, const std::vector<int>& q_lens,
               const std::vector<int>& kv_lens, uint32_t seed, float soft_cap = 0.f, bool alibi = false,
               int window = -1) {
  Problem P;
  P.B = B; P.H = H; P.Hkv = Hkv; P.D = EMU_D; P.bs = bs; P.max_q = max_q;
  P.q_lens = q_lens; P.kv_lens = kv_lens;
  std::mt19937 rng(seed);
  auto frand = [&]() { return (float)((int)(rng() % 65536) - 32768) / 32768.0f; };
  const int D = EMU_D, G = H / Hkv;
  int max_kv = 0, n_blocks = 0;
  P.q_cu = {0}; P.kv_cu = {0}; P.blk_cu = {0};
  for (int b = 0; b < B; ++b) {
    P.q_cu.push_back(P.q_cu.back() + q_lens[b]);
    P.kv_cu.push_back(P.kv_cu.back() + kv_lens[b]);
    const int nb = (kv_lens[b] + bs - 1) / bs;
    P.blk_cu.push_back(P.blk_cu.back() + nb);
    n_blocks += nb;
    max_kv = std::max(max_kv, kv_lens[b]);
  }
  const int pool = n_blocks + 5;
  std::vector<int> ids(pool);
  for (int i = 0; i < pool; ++i) ids[i] = i;
  std::shuffle(ids.begin(), ids.end(), rng);
  for (int i = 0; i < n_blocks; ++i) P.table.push_back(ids[i] * bs);  // first-slot ids
  P.n_slots = (int64_t)pool * bs;
  P.kc.resize((size_t)P.n_slots * Hkv * D);
  P.vc.resize(P.kc.size());
  for (auto& x : P.kc) x = f2bf(frand());
  for (auto& x : P.vc) x = f2bf(frand());
  // slots past a sequence's kv_len inside its last block hold NaN in V and huge K: must be masked
  for (int b = 0; b < B; ++b) {
    const int last = kv_lens[b];
    const int nb = (last + bs - 1) / bs;
    for (int pos = last; pos < nb * bs; ++pos) {
      const int64_t slot = P.table[P.blk_cu[b] + pos / bs] + pos % bs;
      for (int h = 0; h < Hkv; ++h)
        for (int d = 0; d < D; ++d) {
          P.kc[(slot * Hkv + h) * D + d] = f2bf(1e30f);
          P.vc[(slot * Hkv + h) * D + d] = bf16_t{0x7fc0};
        }
    }
  }
  const int T = P.q_cu.back();
  P.q.resize((size_t)T * H * D);
  for (auto& x : P.q) x = f2bf(frand());
  P.out.assign(P.q.size(), f2bf(-77.f));

  int64_t plan[8];
  if (b200_debug_attn_plan(B, max_q, max_kv, H, Hkv, D, bs, plan) != 0 || plan[0] != (EMU_SIMT ? 0 : 2)) {
    std::fprintf(stderr, "attn_emu: the library does not plan the expected kernel for this case (impl %lld)\n",
                 (long long)plan[0]);
    return 1;
  }
  const int n_splits = (int)plan[1], ntm = (int)plan[3];
  const int n_tiles_max = (max_kv + ATT_TILE - 1) / ATT_TILE;
  // fixed-split kernel: any tiles-per-split that covers the tiles with n_splits pieces is valid
  const int tpw = EMU_SIMT ? (n_tiles_max + n_splits - 1) / n_splits : (int)plan[2];
  P.n_seq = (int)plan[4];
  P.total_tiles = plan[6];
  P.ws_o.assign((size_t)B * max_q * H * n_splits * D, NAN);
  P.ws_lse.assign((size_t)B * max_q * H * n_splits, -INFINITY);
  AttnParams& p = P.p;
  std::memset(&p, 0, sizeof(p));
  p.q = P.q.data(); p.out = P.out.data();
  p.q_cu_lens = P.q_cu.data(); p.kv_cu_lens = P.kv_cu.data();
  p.block_table = P.table.data(); p.block_cu_lens = P.blk_cu.data();
  p.ws_o = P.ws_o.data(); p.ws_lse = P.ws_lse.data();
  p.q_stride_t = (int64_t)H * D; p.q_stride_h = D; p.o_stride_t = (int64_t)H * D; p.o_stride_h = D;
  const int R = G >= 4 ? 4 : (G >= 2 ? 2 : 1);  // hg_rows(): query heads per CTA of the CUDA-core kernel
  p.n_heads = H; p.n_kv_heads = Hkv; p.group = G; p.n_hg = EMU_SIMT ? (G + R - 1) / R : 1; p.n_rb = (int)plan[5];
  int shift = 0;
  while ((1 << shift) < bs) ++shift;
  p.block_shift = shift; p.block_mask = bs - 1;
  p.box_rows = bs < ATT_TILE ? bs : ATT_TILE; p.boxes_per_tile = ATT_TILE / p.box_rows;
  p.max_q_len = max_q; p.window = window;
  static std::vector<float> slopes;
  slopes.assign(H, 0.f);
  for (int h = 0; h < H; ++h) slopes[h] = 0.02f * (float)(h + 1);
  p.alibi = alibi ? slopes.data() : nullptr;
  const float sm_scale = 1.0f / std::sqrt((float)D);
  constexpr float LOG2E = 1.4426950408889634f;
  if (soft_cap > 0.f) {  // as b200_paged_attn_decode sets them
    p.use_cap = 1;
    p.cap_in = sm_scale / soft_cap;
    p.cap_out_log2 = soft_cap * LOG2E;
  } else {
    p.use_cap = 0;
    p.scale_log2 = sm_scale * LOG2E;
  }
  p.n_splits = n_splits; p.tiles_per_split = tpw; p.ntm = ntm; p.tpw = tpw; p.stream = EMU_SIMT ? 0 : 1;
  P.kmap = CUtensorMap{P.kc.data(), P.n_slots, Hkv, D, p.box_rows};
  P.vmap = CUtensorMap{P.vc.data(), P.n_slots, Hkv, D, p.box_rows};
  g_pr = &P;

#if EMU_SIMT
  unsigned grid = 0;
  for (unsigned bz = 0; bz < (unsigned)(B * max_q) && !g_failed; ++bz)
    for (unsigned by = 0; by < (unsigned)(Hkv * p.n_hg); ++by)
      for (unsigned bx = 0; bx < (unsigned)n_splits; ++bx, ++grid) {
        g_bar_base = nullptr;
        for (auto& b : g_eb) b = EmuBar{};
        for (auto& w : g_warps) pthread_barrier_init(&w.bar, nullptr, LANES);
        pthread_barrier_init(&g_cta_bar, nullptr, 4 * LANES);
        pthread_t th[4 * LANES];
        struct SArg { int tid; unsigned bx, by, bz; int R; };
        static SArg sargs[4 * LANES];
        for (int i = 0; i < 4 * LANES; ++i) {
          sargs[i] = SArg{i, bx, by, bz, R};
          pthread_create(&th[i], nullptr, [](void* a) -> void* {
            SArg* x = (SArg*)a;
            t_lane = x->tid & 31;
            t_warp = &g_warps[x->tid >> 5];
            threadIdx.x = (unsigned)x->tid;
            blockIdx.x = x->bx; blockIdx.y = x->by; blockIdx.z = x->bz;
            Problem& Q = *g_pr;
            if (x->R == 4) paged_attn_decode_kernel<bf16_t, EMU_D, 4>(Q.kmap, Q.vmap, Q.p);
            else if (x->R == 2) paged_attn_decode_kernel<bf16_t, EMU_D, 2>(Q.kmap, Q.vmap, Q.p);
            else paged_attn_decode_kernel<bf16_t, EMU_D, 1>(Q.kmap, Q.vmap, Q.p);
            return nullptr;
          }, &sargs[i]);
        }
        for (int i = 0; i < 4 * LANES; ++i) pthread_join(th[i], nullptr);
        for (auto& w : g_warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&g_cta_bar);
      }
#else
  const unsigned grid = (unsigned)((P.total_tiles + tpw - 1) / tpw);
  for (unsigned cta = 0; cta < grid && !g_failed; ++cta) {
    g_bar_base = nullptr;
    for (auto& b : g_eb) b = EmuBar{};
    pthread_barrier_init(&g_warps[0].bar, nullptr, LANES);
    pthread_t th[LANES];
    struct Arg { int lane; unsigned cta; };
    static Arg args[LANES];
    for (int i = 0; i < LANES; ++i) {
      args[i] = Arg{i, cta};
      pthread_create(&th[i], nullptr, [](void* a) -> void* {
        Arg* x = (Arg*)a;
        blockIdx.x = x->cta;
        t_warp = &g_warps[0];
        return lane_main((void*)(intptr_t)x->lane);
      }, &args[i]);
    }
    for (int i = 0; i < LANES; ++i) pthread_join(th[i], nullptr);
    pthread_barrier_destroy(&g_warps[0].bar);
  }
#endif
  if (g_failed) return 1;

  // second pass, as launch_attn does it: grid ((n_heads + 3) / 4, batch * max_q_len), 4 warps per
  // CTA, each warp on its own (token, head) row: the warps are run one after the other
  if (n_splits > 1) {
    for (unsigned by = 0; by < (unsigned)(B * max_q) && !g_failed; ++by)
      for (unsigned bx = 0; bx < (unsigned)((H + 3) / 4); ++bx)
        for (int w = 0; w < 4; ++w) {
          pthread_barrier_init(&g_warps[0].bar, nullptr, LANES);
          pthread_t th[LANES];
          struct CArg { int lane, warp; unsigned bx, by; };
          static CArg cargs[LANES];
          for (int i = 0; i < LANES; ++i) {
            cargs[i] = CArg{i, w, bx, by};
            pthread_create(&th[i], nullptr, [](void* a) -> void* {
              CArg* x = (CArg*)a;
              t_lane = x->lane;
              t_warp = &g_warps[0];
              threadIdx.x = (unsigned)(x->warp * 32 + x->lane);
              blockIdx.x = x->bx;
              blockIdx.y = x->by;
              paged_attn_combine_kernel<bf16_t, EMU_D>(g_pr->p);
              return nullptr;
            }, &cargs[i]);
          }
          for (int i = 0; i < LANES; ++i) pthread_join(th[i], nullptr);
          pthread_barrier_destroy(&g_warps[0].bar);
        }
  }

  // compare the kernel pair's output with the reference (and, for split KV, also the harness's
  // own LSE merge of the partials)
  double worst = 0.0;
  for (int b = 0; b < B; ++b)
    for (int qi = 0; qi < q_lens[b]; ++qi)
      for (int h = 0; h < H; ++h) {
        const int kvh = h / G, end = kv_lens[b] - q_lens[b] + qi + 1;
        const int64_t tok = P.q_cu[b] + qi;
        std::vector<double> s(end);
        double mx = -1e300;
        const int begin = window >= 0 ? std::max(0, end - 1 - window) : 0;
        for (int j = 0; j < end; ++j) {
          const int64_t slot = P.table[P.blk_cu[b] + j / bs] + j % bs;
          double a = 0;
          for (int d = 0; d < D; ++d)
            a += (double)bf2f(P.q[(tok * H + h) * D + d]) * bf2f(P.kc[(slot * Hkv + kvh) * D + d]);
          a /= std::sqrt((double)D);
          if (soft_cap > 0.f) a = soft_cap * std::tanh(a / soft_cap);
          if (alibi) a += (double)slopes[h] * j;
          s[j] = j >= begin ? a : -1e300;
          mx = std::max(mx, s[j]);
        }
        double sum = 0;
        for (int j = 0; j < end; ++j) sum += (s[j] = j >= begin ? std::exp(s[j] - mx) : 0.0);
        const int64_t wrow = ((int64_t)b * max_q + qi) * H + h;
        double M = -INFINITY;
        for (int sp = 0; sp < n_splits; ++sp) M = std::max(M, (double)P.ws_lse[wrow * n_splits + sp]);
        for (int d = 0; d < D; ++d) {
          double want = 0;
          for (int j = 0; j < end; ++j) {
            const int64_t slot = P.table[P.blk_cu[b] + j / bs] + j % bs;
            want += s[j] * bf2f(P.vc[(slot * Hkv + kvh) * D + d]);
          }
          want /= sum;
          double got = bf2f(P.out[(tok * H + h) * D + d]);  // stream kernel (+ combine pass)
          {
            const double err0 = std::fabs(got - want);
            if (!(err0 <= worst)) worst = err0;
          }
          if (n_splits > 1) {
            double L = 0, O = 0;
            for (int sp = 0; sp < n_splits; ++sp) {
              const double lse = P.ws_lse[wrow * n_splits + sp];
              if (lse == -INFINITY) continue;
              const double w = std::exp2(lse - M);
              L += w;
              O += w * P.ws_o[(wrow * n_splits + sp) * D + d];
            }
            got = O / L;
          }
          const double err = std::fabs(got - want);
          if (!(err <= worst)) worst = err;
        }
      }
  std::printf("B=%d H=%d/%d bs=%d max_q=%d max_kv=%d: grid %u, tpw %d, n_splits %d, max |err| %.5f\n", B, H, Hkv, bs,
              max_q, max_kv, grid, tpw, n_splits, worst);
  return worst < 2e-2 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(3, 8, 2, 8, 1, {1, 1, 1}, {127, 300, 40}, 1);         // GQA 4, decode, ragged lengths
  bad += run(2, 4, 4, 16, 1, {1, 1}, {33, 257}, 2);                 // MHA, block 16
  bad += run(2, 8, 1, 1, 1, {1, 1}, {70, 19}, 3);                   // MQA 8 rows, block size 1
  if (EMU_SIMT) bad += run(2, 6, 3, 8, 2, {2, 1}, {100, 37}, 6);   // group 2 (two query heads per CTA)
  if (!EMU_TR) bad += run(2, 8, 2, 8, 3, {3, 2}, {100, 37}, 4);     // multi-token queries (12 rows: not a TR shape)
  else bad += run(2, 8, 2, 8, 2, {2, 1}, {100, 37}, 4);             // 8 packed rows, causal diagonal inside
  bad += run(1, 8, 2, 8, 1, {1}, {1500}, 5);                        // long: many pieces over many warps
  bad += run(2, 8, 2, 8, 1, {1, 1}, {200, 90}, 7, 30.f, true, 40);  // soft cap + alibi + sliding window
  bad += run(2, 4, 4, 16, 1, {1, 1}, {130, 64}, 8, 0.f, false, 0);  // window 0: only the token itself
  std::printf(bad || g_failed ? "FAILED\n" : "ok\n");
  return bad || g_failed ? 1 : 0;
}
