// attn_tr_emu.cpp — host execution of the transposed attention tile AS WRITTEN in
// scalellm_b200/csrc/paged_attn.cu (TR = 1 branches of paged_attn_persist_kernel).
//
// tools/attn_tr_emu.py cuts the three marked blocks ([tr-emu:load_q], [tr-emu:tile],
// [tr-emu:finalize]) out of the .cu file into attn_tr_emu_*.inc and compiles this harness around
// them (-DEMU_TR=1).  With -DEMU_TR=0 the default branches ([def-emu:*]) are compiled instead: they
// are validated on the GPU, so running them here validates the emulated ldmatrix / mma.sync /
// shuffle semantics the TR check relies on.  How it works: 32 host threads play the lanes of one
// warp and run the kernel's own statements; the warp
// collectives (ldmatrix, mma.sync, movmatrix, shuffles, votes) are emulated with a barrier and a
// shared exchange area, shared memory is a byte array filled the way TMA fills it (128-byte
// swizzle).  The result is compared with a plain softmax(QK^T)V.  This checks the CUDA source's
// index arithmetic (fragment ownership, swizzled addresses, masks, output scatter) without a GPU;
// tools/attn_tr_model.py checks the algebra the source was written from.
#include <pthread.h>

#ifndef EMU_TR
#define EMU_TR 1
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

// ---- element type stand-in -------------------------------------------------------------------
struct bf16_t {
  uint16_t bits;
};
static inline float bf2f(bf16_t x) {
  uint32_t u = (uint32_t)x.bits << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline bf16_t f2bf(float f) {  // round to nearest even
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
};

// ---- the warp ------------------------------------------------------------------------------------
constexpr int LANES = 32;
static pthread_barrier_t g_bar;
static thread_local int t_lane;
static uint8_t g_smem[64 * 1024] __attribute__((aligned(1024)));
static uint32_t g_x32[LANES][8];  // exchange area of the collectives
static float g_xf[LANES];
static int g_xi[LANES];
static void sync_warp() { pthread_barrier_wait(&g_bar); }

static inline float __shfl_xor_sync(unsigned, float v, int mask) {
  g_xf[t_lane] = v;
  sync_warp();
  const float r = g_xf[t_lane ^ mask];
  sync_warp();
  return r;
}
static inline bool __any_sync(unsigned, bool pred) {
  g_xi[t_lane] = pred;
  sync_warp();
  bool r = false;
  for (int i = 0; i < LANES; ++i) r |= g_xi[i] != 0;
  sync_warp();
  return r;
}
static inline uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }

// ldmatrix .x4: lanes 8j..8j+7 give the 16-byte row addresses of matrix j; lane L receives, of
// each matrix, row L/4 columns 2(L%4), +1 (or, transposed, rows 2(L%4), +1 of column L/4)
static void ldsm_impl(uint32_t (&r)[4], uint32_t addr, bool trans) {
  g_x32[t_lane][0] = addr;
  sync_warp();
  for (int j = 0; j < 4; ++j) {
    uint16_t e[2];
    for (int k = 0; k < 2; ++k) {
      const int row = trans ? 2 * (t_lane & 3) + k : t_lane >> 2;
      const int col = trans ? t_lane >> 2 : 2 * (t_lane & 3) + k;
      std::memcpy(&e[k], g_smem + g_x32[8 * j + row][0] + col * 2, 2);
    }
    r[j] = (uint32_t)e[0] | ((uint32_t)e[1] << 16);
  }
  sync_warp();
}
static void ldsm_x4(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, false); }
static void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, true); }

static uint32_t movmatrix_trans(uint32_t a) {
  g_x32[t_lane][0] = a;
  sync_warp();
  uint16_t e[2];
  for (int k = 0; k < 2; ++k) {  // out(row L/4, col 2(L%4)+k) = in(row 2(L%4)+k, col L/4)
    const int src_row = 2 * (t_lane & 3) + k, src_col = t_lane >> 2;
    const uint32_t w = g_x32[src_row * 4 + src_col / 2][0];
    e[k] = (uint16_t)(src_col & 1 ? w >> 16 : w & 0xffffu);
  }
  sync_warp();
  return (uint32_t)e[0] | ((uint32_t)e[1] << 16);
}

// mma.sync.m16n8k16 row.col, bf16 inputs, fp32 accumulate
template <typename T>
static void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  for (int i = 0; i < 4; ++i) g_x32[t_lane][i] = a[i];
  g_x32[t_lane][4] = b0;
  g_x32[t_lane][5] = b1;
  sync_warp();
  auto half = [](uint32_t w, int k) { return bf2f(bf16_t{(uint16_t)(k ? w >> 16 : w & 0xffffu)}); };
  auto A = [&](int row, int col) {  // a0: (g, 2t..) a1: (g+8, 2t..) a2: (g, 2t+8..) a3: (g+8, 2t+8..)
    const int g = row & 7, t = (col & 7) >> 1;
    return half(g_x32[g * 4 + t][(row >> 3) + 2 * (col >> 3)], col & 1);
  };
  auto B = [&](int k, int n) {  // b0: (k = 2t.., n = g)  b1: (k = 2t+8.., n = g)
    return half(g_x32[n * 4 + ((k & 7) >> 1)][4 + (k >> 3)], k & 1);
  };
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

// ---- the kernel's surroundings -------------------------------------------------------------------
struct Params {
  const void* q;
  void* out;
  const float* alibi;
  float* ws_o;
  float* ws_lse;
  int64_t q_stride_t, q_stride_h, o_stride_t, o_stride_h;
  int n_heads, n_splits, max_q_len, use_cap;
  float scale_log2, cap_in, cap_out_log2;
};
struct Item {
  int n_tiles, q_len, rb, kvh, q_begin, b, split;
};

constexpr int D = 128, ATT_TILE = 16;
using T = bf16_t;

struct Case {
  int G, q_len, kv_len, n_tiles;
  Params p;
  std::vector<bf16_t> q, out;      // q: [q_len][G heads][D]
  std::vector<bf16_t> k, v;        // [n_tiles * 16][D] (row = key position)
  std::vector<float> ws_o, ws_lse;
};
static Case* g_case;

static void* lane_main(void* arg) {
  t_lane = (int)(intptr_t)arg;
  const int lane = t_lane;
  Case& c = *g_case;
  const Params& p = c.p;
  constexpr int KS = D / 16, NB = EMU_TR ? D / 16 : D / 8, QR = EMU_TR ? 2 : 4, ROWS = EMU_TR ? 8 : 16;
  constexpr int ROWB = D * (int)sizeof(T);
  const int G = c.G;
  Item cur{c.n_tiles, c.q_len, 0, 0, 0, 0, 0};
  const Item& it = cur;
  uint32_t qa[KS][QR];
  const int rows_total = it.q_len * G, row0 = it.rb * ROWS;
  const int n_rows = rows_total - row0 < ROWS ? rows_total - row0 : ROWS;
  {
#include "attn_tr_emu_load_q.inc"
  }
  // per-item row bookkeeping, as in the kernel (window / alibi off)
  const int q_pos0 = c.kv_len - cur.q_len;
  int row_end[2], row_begin[2];
  float slope_log2[2] = {0.f, 0.f};
  for (int h = 0; h < 2; ++h) {
    const int r = EMU_TR ? (lane & 3) * 2 + h : (lane >> 2) + 8 * h;
    const bool ok = r < n_rows;
    const int row = row0 + (ok ? r : 0), qi = row / G;
    row_end[h] = (ok || EMU_TR) ? q_pos0 + qi + 1 : 0;
    row_begin[h] = 0;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[NB][4];
  for (int nb = 0; nb < NB; ++nb)
    for (int e = 0; e < 4; ++e) o[nb][e] = 0.f;
  const int lm = lane >> 3, lr = lane & 7;
  const int kv_end = q_pos0 + (row0 + n_rows - 1) / G + 1;
  for (int i = 0; i < cur.n_tiles; ++i) {
    // "TMA": tile i of K and V into shared memory with the 128-byte swizzle; rows past the causal
    // end hold garbage in K and zeros in V (the kernel zeroes V rows itself on a boundary tile)
    const uint32_t k_base = 0, v_base = ATT_TILE * ROWB;
    if (lane == 0) {
      for (int r = 0; r < ATT_TILE; ++r)
        for (int d = 0; d < D; ++d) {
          const int pos = i * ATT_TILE + r;
          bf16_t kk = c.k[(size_t)pos * D + d], vv = c.v[(size_t)pos * D + d];
          if (pos >= kv_end) vv = f2bf(0.f);
          std::memcpy(g_smem + swz128(k_base + r * ROWB + d * 2), &kk, 2);
          std::memcpy(g_smem + swz128(v_base + r * ROWB + d * 2), &vv, 2);
        }
    }
    sync_warp();
    const int pos0 = i * ATT_TILE;
    {
#include "attn_tr_emu_tile.inc"
    }
    sync_warp();
  }
  {
#include "attn_tr_emu_finalize.inc"
  }
  return nullptr;
}

static float frand(uint32_t& s) {
  s = s * 1664525u + 1013904223u;
  return ((s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

static int run(int G, int q_len, int kv_len, int n_splits, uint32_t seed) {
  Case c;
  c.G = G;
  c.q_len = q_len;
  c.kv_len = kv_len;
  c.n_tiles = (kv_len + ATT_TILE - 1) / ATT_TILE;
  const int H = G;  // one kv head
  c.q.resize((size_t)q_len * H * D);
  c.out.assign((size_t)q_len * H * D, f2bf(-77.f));
  c.k.resize((size_t)c.n_tiles * ATT_TILE * D);
  c.v.resize(c.k.size());
  for (auto& x : c.q) x = f2bf(frand(seed));
  for (size_t i = 0; i < c.k.size(); ++i) {
    const bool past = (int)(i / D) >= kv_len;
    c.k[i] = f2bf(past ? 1e30f : frand(seed));     // stale keys must be masked, not multiplied
    c.v[i] = past ? bf16_t{0x7fc0} : f2bf(frand(seed));  // NaN past the end of the sequence
  }
  c.ws_o.assign((size_t)q_len * H * n_splits * D, -55.f);
  c.ws_lse.assign((size_t)q_len * H * n_splits, -55.f);
  c.p = Params{c.q.data(), c.out.data(), nullptr, c.ws_o.data(), c.ws_lse.data(), (int64_t)H * D, D, (int64_t)H * D, D,
               H, n_splits, q_len, 0, 1.4426950408889634f / std::sqrt((float)D), 0.f, 0.f};
  g_case = &c;
  pthread_barrier_init(&g_bar, nullptr, LANES);
  pthread_t th[LANES];
  for (int i = 0; i < LANES; ++i) pthread_create(&th[i], nullptr, lane_main, (void*)(intptr_t)i);
  for (int i = 0; i < LANES; ++i) pthread_join(th[i], nullptr);
  pthread_barrier_destroy(&g_bar);

  // reference: softmax(q k^T / sqrt(D)) v per (token, head), causal with diagonal kv_len - q_len
  double worst = 0.0;
  for (int qi = 0; qi < q_len; ++qi)
    for (int h = 0; h < H; ++h) {
      const int end = kv_len - q_len + qi + 1;
      std::vector<double> s(end);
      double mx = -1e300;
      for (int j = 0; j < end; ++j) {
        double a = 0;
        for (int d = 0; d < D; ++d) a += (double)bf2f(c.q[((size_t)qi * H + h) * D + d]) * bf2f(c.k[(size_t)j * D + d]);
        s[j] = a / std::sqrt((double)D);
        mx = s[j] > mx ? s[j] : mx;
      }
      double sum = 0;
      for (int j = 0; j < end; ++j) sum += (s[j] = std::exp(s[j] - mx));
      for (int d = 0; d < D; ++d) {
        double oo = 0;
        for (int j = 0; j < end; ++j) oo += s[j] * bf2f(c.v[(size_t)j * D + d]);
        oo /= sum;
        const double got = n_splits == 1 ? bf2f(c.out[((size_t)qi * H + h) * D + d])
                                         : c.ws_o[(((size_t)qi * H + h) * n_splits + 0) * D + d];
        const double err = std::fabs(got - oo);
        if (!(err <= worst)) worst = err;  // NaN-propagating max
      }
      if (n_splits > 1) {
        const double lse2 = (mx + std::log(sum)) * 1.4426950408889634;
        const double err = std::fabs(c.ws_lse[((size_t)qi * H + h) * n_splits] - lse2);
        if (!(err <= 1e-3)) worst = 1e9;
      }
    }
  std::printf("G=%d q_len=%d kv_len=%d n_splits=%d: max |err| %.5f\n", G, q_len, kv_len, n_splits, worst);
  return worst < 2e-2 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(4, 1, 80, 1, 1);    // the benchmark's shape of rows: 4 of 8
  bad += run(4, 1, 53, 2, 2);    // ragged end, split-KV partial + LSE
  bad += run(8, 1, 33, 1, 3);    // all 8 rows
  bad += run(4, 2, 47, 1, 4);    // two query tokens: causal diagonal inside the block
  bad += run(1, 5, 21, 2, 5);    // MHA rows = tokens
  std::printf(bad ? "FAILED\n" : "ok\n");
  return bad ? 1 : 0;
}
