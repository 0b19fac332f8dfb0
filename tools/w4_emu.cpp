// w4_emu.cpp — host execution of the W4A16 GEMM kernel's role code AS WRITTEN in
// scalellm_b200/csrc/w4a16.cu, for one CTA.
//
// tools/w4_emu.py cuts the marked blocks out of the sources ([w4-emu:plan] from common.cuh,
// [w4-emu:cfg], [w4-emu:init], [w4-emu:roles] from w4a16.cu) and compiles this harness around them
// once per kernel variant (-DEMU_VAR=...).  One host thread per warp runs the kernel's own
// statements (as lane 0); mbarriers, the bulk / tensor copies, tensor memory and the in-order
// tensor pipe are emulated.  Data is replaced by tile ids, so what is checked is the protocol the
// source implements: every MMA reads a TMEM slot and an activation stage that hold the tile it is
// meant for, nothing is overwritten under a queued MMA, every accumulator segment holds exactly
// its k range when it is committed and is drained before it is reused, and nothing deadlocks.
// (tools/w4_protocol_sim.py checks the same rules on a hand-written model of the protocol.)
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#ifndef EMU_VAR
#define EMU_VAR 0
#endif
#ifndef EMU_MT
#define EMU_MT 64
#endif

// ---- CUDA-isms the blocks use ----------------------------------------------------------------------
#define __host__
#define __device__
#define __forceinline__ inline
#define __restrict__
struct uint4 {
  uint32_t x, y, z, w;
};
struct __nv_bfloat162 {
  uint32_t bits;
};
struct CUtensorMap {
  int unused;
};
struct Dim3 {
  int x;
};
static thread_local Dim3 threadIdx, blockIdx;
static inline long long clock64() { return 0; }
static inline float __uint_as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
using std::min;
constexpr int W4_QBYTES = 8192;
constexpr int W4_MAX_BLOB = 8192 + 4 * (256 + 128);

#include "w4_emu_plan.inc"

constexpr int MT = EMU_MT, NSUB = 1, VAR = EMU_VAR;
constexpr bool TRACE = false;
#include "w4_emu_cfg.inc"
#undef W4_TRACE
#undef W4_TRACE_NS
#define W4_TRACE(slot) do { } while (0)
#define W4_TRACE_NS(slot) do { } while (0)
using Cfg = W4Cfg<MT, NSUB>;

// ---- failure reporting -----------------------------------------------------------------------------
static std::atomic<bool> g_failed{false};
static std::mutex g_mu;  // guards every piece of emulated hardware state
static void fail(const char* what, long a = 0, long b = 0, long c = 0) {
  if (!g_failed.exchange(true)) std::fprintf(stderr, "w4_emu[VAR=%d]: %s (%ld %ld %ld)\n", VAR, what, a, b, c);
}

// ---- shared memory, barriers -------------------------------------------------------------------------
static uint8_t g_smem[Cfg::SMEM + 4096] __attribute__((aligned(1024)));
static uint8_t* const act_smem = g_smem;
static uint8_t* const raw_smem = act_smem + Cfg::ACT_STAGES * Cfg::ACT_BYTES;
static uint64_t g_bars[Cfg::N_BARS];
static uint64_t* const raw_full = g_bars;
static uint64_t* const raw_empty = raw_full + Cfg::RAW_STAGES;
static uint64_t* const act_full = raw_empty + Cfg::RAW_STAGES;
static uint64_t* const act_empty = act_full + Cfg::ACT_STAGES;
static uint64_t* const deq_full = act_empty + Cfg::ACT_STAGES;
static uint64_t* const deq_empty = deq_full + Cfg::A_STAGES;
static uint64_t* const tmem_full = deq_empty + Cfg::A_STAGES;
static uint64_t* const tmem_empty = tmem_full + 2;
static uint32_t g_holder[16];  // the kernel's tmem_holder word and what follows it in shared memory
static uint32_t* const tmem_holder = g_holder;

struct EmuBar {
  int count = 0, pending = 0, phase = 0;
  long tx = 0;
};
static EmuBar g_eb[Cfg::N_BARS];
static bool g_drained[2] = {true, true};
static bool g_readable[2] = {false, false};  // accumulator holds (part of) the segment handed to the epilogue
static int g_reads[2] = {0, 0};              // tcgen05.ld of it since the hand-over
static inline int bar_index(const uint64_t* b) { return (int)(b - g_bars); }
static void segment_complete(int buf);  // an accumulator segment was handed to the epilogue
static void bar_check_complete(int i) {  // g_mu held
  EmuBar& b = g_eb[i];
  if (b.pending == 0 && b.tx == 0) {
    ++b.phase;
    b.pending = b.count;
    const int te = bar_index(tmem_empty), tf = bar_index(tmem_full);
    if (i == te || i == te + 1) {  // the epilogue released the accumulator(s)
      for (int a = 0; a < 2; ++a) {
        if (!(VAR & 16) && a != i - te) continue;
        if (g_readable[a] && g_reads[a] == 0) fail("epilogue released an accumulator it never read", a);
        g_readable[a] = false;
        g_drained[a] = true;
      }
    }
    if (i == tf || i == tf + 1) segment_complete(i - tf);
  }
}
static void mbar_init(uint64_t* bar, uint32_t count) {
  EmuBar& b = g_eb[bar_index(bar)];
  b.count = b.pending = (int)count;
  b.phase = 0;
  b.tx = 0;
}
static void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> lk(g_mu);
  EmuBar& b = g_eb[bar_index(bar)];
  if (--b.pending < 0) fail("more arrivals than the barrier expects", bar_index(bar));
  bar_check_complete(bar_index(bar));
}
static void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  EmuBar& b = g_eb[bar_index(bar)];
  b.tx += bytes;
  if (--b.pending < 0) fail("more arrivals than the barrier expects", bar_index(bar));
  bar_check_complete(bar_index(bar));
}
static void bar_complete_tx(int i, long bytes) {  // g_mu held
  g_eb[i].tx -= bytes;
  bar_check_complete(i);
}
static void mbar_wait(uint64_t* bar, uint32_t parity) {
  const int i = bar_index(bar);
  for (int spin = 0;; ++spin) {
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if ((uint32_t)(g_eb[i].phase & 1) != parity) return;
    }
    if (g_failed.load()) return;  // let the roles run off the end so the process can report
    if (spin < 200) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(50));  // do not starve the other roles
  }
}

// ---- no-ops and trivia ---------------------------------------------------------------------------------
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tmem_st_wait() {}
static inline void tmem_ld_wait() {}
static inline void __syncwarp() {}
static inline void pdl_wait() {}
static inline void pdl_launch_dependents() {}
static inline bool elect_one() { return true; }
template <typename V>
static inline V __shfl_sync(unsigned, V v, int) { return v; }
static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - g_smem); }
static constexpr uint32_t umma_idesc_bf16(uint32_t, uint32_t) { return 0; }
static inline uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) { return smem_addr >> 4; }

// ---- copies: the weight-blob ring and the activation ring ---------------------------------------------
struct Copy {
  int kind;  // 0: weight blob, 1: activation half
  int stage, half;
  long tile;  // weight tile index / k tile
  int bar;
  long bytes;
};
static std::vector<Copy> g_copies;            // in flight (g_mu)
static long g_raw_content[Cfg::RAW_STAGES];   // weight tile index held by each ring entry
static long g_act_content[Cfg::ACT_STAGES][2];  // k tile held by each activation stage half

struct PipeOp {
  int kind;  // 0: mma, 1: commit
  int slot, ks, stage, buf, accumulate, bar;
};
static std::deque<PipeOp> g_pipe;  // the tensor pipe's in-order queue (g_mu)
static long g_slot_content[Cfg::A_STAGES][4][2];  // [slot][lane quadrant][half] -> weight tile index

struct W4ParamsEmu;  // (W4Params comes from the cfg block)
static const uint8_t* const PACKED_BASE = reinterpret_cast<const uint8_t*>(0x100000);

static void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  std::lock_guard<std::mutex> lk(g_mu);
  Copy c{0, (int)(((uint8_t*)smem_dst - raw_smem) / Cfg::RAW_BYTES), 0,
         (long)(((const uint8_t*)gmem_src - PACKED_BASE) / (long)bytes), bar_index(bar), (long)bytes};
  if (c.stage < 0 || c.stage >= Cfg::RAW_STAGES) fail("weight copy outside the ring", c.stage);
  g_copies.push_back(c);
}
static void tma_load_2d(void* smem_dst, const CUtensorMap*, uint64_t* bar, int c0, int c1) {
  std::lock_guard<std::mutex> lk(g_mu);
  const long off = (uint8_t*)smem_dst - act_smem;
  Copy c{1, (int)(off / Cfg::ACT_BYTES), (int)((off % Cfg::ACT_BYTES) / Cfg::ACT_ATOM), c0 / 128, bar_index(bar),
         (long)Cfg::ACT_ATOM};
  if (c1 != 0 || c0 % 64 || (c0 % 128) / 64 != c.half) fail("activation copy coordinates", c0, c1, c.half);
  for (const PipeOp& op : g_pipe)
    if (op.kind == 0 && op.stage == c.stage) fail("activation stage reloaded under a queued MMA", c.stage);
  g_copies.push_back(c);
}
static inline uint4 lds128(uint32_t addr) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int stage = (int)((addr - smem_u32(raw_smem)) / Cfg::RAW_BYTES);
  const uint32_t t = (uint32_t)g_raw_content[stage];
  return uint4{t, t, t, t};
}
static inline uint32_t lds_u16(uint32_t) { return 0; }
static inline uint32_t lds_u8(uint32_t) { return 0; }
static inline __nv_bfloat162 w4_zmagic(uint32_t) { return __nv_bfloat162{0}; }
static inline uint4 w4_dequant_word(uint32_t word, __nv_bfloat162, __nv_bfloat162) { return uint4{word, word, word, word}; }

// ---- tensor memory + tensor pipe -----------------------------------------------------------------------
static void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int col = (int)(taddr & 0xffff) - Cfg::A_COL0, quad = (int)(taddr >> 16) / 32;
  const int slot = col / 64, half = (col % 64) / 32;
  if (col < 0 || slot >= Cfg::A_STAGES || col % 32 || quad < 0 || quad > 3) fail("tcgen05.st address", taddr);
  for (const PipeOp& op : g_pipe)
    if (op.kind == 0 && op.slot == slot) fail("TMEM slot rewritten under a queued MMA", slot);
  for (int i = 1; i < 32; ++i)
    if (r[i] != r[0]) fail("dequantised words of one store come from different tiles", r[0], r[i]);
  g_slot_content[slot][quad][half] = (long)r[0];
}
static void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t, uint32_t accumulate) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int acol = (int)(a_tmem & 0xffff) - Cfg::A_COL0;
  const long boff = (long)(b_desc << 4) - (long)smem_u32(act_smem);
  PipeOp op{0, acol / 64, (acol % 64) / 8, (int)(boff / Cfg::ACT_BYTES), (int)(d_tmem & 0xffff) / MT, (int)accumulate, 0};
  const long in_stage = boff % Cfg::ACT_BYTES;
  const int ks_b = (int)(in_stage / Cfg::ACT_ATOM) * 4 + (int)((in_stage % Cfg::ACT_ATOM) / 32);
  if (acol < 0 || acol % 8 || op.slot >= Cfg::A_STAGES || boff < 0 || op.stage >= Cfg::ACT_STAGES ||
      (in_stage % Cfg::ACT_ATOM) % 32 || ks_b != op.ks || (d_tmem & 0xffff) % MT || op.buf > 1)
    fail("tcgen05.mma operands", a_tmem, (long)b_desc, d_tmem);
  g_pipe.push_back(op);
}
static void umma_commit(uint64_t* bar) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_pipe.push_back(PipeOp{1, 0, 0, 0, 0, 0, bar_index(bar)});
}
static void tmem_ld_check(uint32_t taddr) {  // the epilogue may only read what was handed over to it
  std::lock_guard<std::mutex> lk(g_mu);
  const int a = (int)(taddr & 0xffff) / MT;
  if (a < 0 || a > 1 || !g_readable[a]) fail("epilogue reads an accumulator that holds no part of its segment", a);
  else ++g_reads[a];
}
static inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_check(taddr); for (auto& x : r) x = 0; }
static inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_check(taddr); for (auto& x : r) x = 0; }

// ---- expected work of this CTA ---------------------------------------------------------------------------
struct Segment {
  int nt, kt0, kt1;
};
static std::vector<Segment> g_segments;
static int g_KT = 0;
static size_t g_seg_committed = 0;
static std::vector<long> g_acc[2];  // weight tiles accumulated into each buffer since its last reset
static int g_acc_ks[2] = {0, 0};    // k-steps of the tile being accumulated

static bool g_fresh[2] = {false, false};  // accumulator was (re)started since the last segment hand-over

static void segment_complete(int buf) {  // g_mu held: tmem_full[buf] completed a phase
  if (g_seg_committed >= g_segments.size()) { fail("more accumulator hand-overs than segments"); return; }
  const Segment& s = g_segments[g_seg_committed];
  long cnt0 = 0;  // tiles of the CTA before this segment
  for (size_t i = 0; i < g_seg_committed; ++i) cnt0 += g_segments[i].kt1 - g_segments[i].kt0;
  if (VAR & 16) {  // two issuers: tiles with even / odd position in the CTA's tile stream
    if (buf != 0) fail("two-issuer kernel uses tmem_full[0] only", buf);
    std::vector<long> want[2];
    for (int kt = s.kt0; kt < s.kt1; ++kt) want[(cnt0 + kt - s.kt0) & 1].push_back((long)s.nt * g_KT + kt);
    for (int i = 0; i < 2; ++i) {
      if (want[i].empty() ? g_fresh[i] : (!g_fresh[i] || g_acc[i] != want[i]) || g_acc_ks[i] != 0)
        fail("issuer's accumulator does not hold exactly its share of the segment", (long)g_seg_committed, i,
             (long)g_acc[i].size());
      g_fresh[i] = false;
      g_readable[i] = !want[i].empty();
      g_reads[i] = 0;
    }
  } else {
    if ((int)(g_seg_committed & 1) != buf) fail("segment committed from the wrong accumulator buffer", buf);
    std::vector<long> want;
    for (int kt = s.kt0; kt < s.kt1; ++kt) want.push_back((long)s.nt * g_KT + kt);
    if (g_acc[buf] != want || g_acc_ks[buf] != 0)
      fail("accumulator does not hold exactly its segment's tiles", (long)g_seg_committed, (long)g_acc[buf].size(),
           (long)want.size());
    g_readable[buf] = true;
    g_reads[buf] = 0;
  }
  ++g_seg_committed;
}

static void pipe_execute(const PipeOp& op) {  // g_mu held
  if (op.kind == 1) {
    if (--g_eb[op.bar].pending < 0) fail("more arrivals than the barrier expects", op.bar);
    bar_check_complete(op.bar);
    return;
  }
  long tile = g_slot_content[op.slot][0][0];
  for (int q = 0; q < 4; ++q)
    for (int h = 0; h < 2; ++h)
      if (g_slot_content[op.slot][q][h] != tile) {
        if (!g_failed.load()) {
          std::fprintf(stderr, "  slot %d contents:", op.slot);
          for (int qq = 0; qq < 4; ++qq) std::fprintf(stderr, " [%ld %ld]", g_slot_content[op.slot][qq][0], g_slot_content[op.slot][qq][1]);
          std::fprintf(stderr, "  act stage %d: %ld %ld  ks %d buf %d acc %d\n", op.stage, g_act_content[op.stage][0],
                       g_act_content[op.stage][1], op.ks, op.buf, op.accumulate);
          const int df = bar_index(deq_full) + op.slot, de = bar_index(deq_empty) + op.slot;
          std::fprintf(stderr, "  deq_full: phase %d pending %d tx %ld   deq_empty: phase %d pending %d\n", g_eb[df].phase,
                       g_eb[df].pending, g_eb[df].tx, g_eb[de].phase, g_eb[de].pending);
        }
        fail("TMEM slot holds pieces of different tiles", op.slot);
      }
  const long kt = g_act_content[op.stage][0];
  if (g_act_content[op.stage][1] != kt) fail("activation stage halves hold different k tiles", op.stage);
  if (tile < 0 || kt < 0 || tile % g_KT != kt) fail("MMA pairs a weight tile with another k tile's activations", tile, kt);
  std::vector<long>& acc = g_acc[op.buf];
  if (op.ks == 0) {
    if (!op.accumulate) {  // first MMA of a segment overwrites the accumulator
      if (!g_drained[op.buf]) fail("accumulator overwritten before the epilogue drained it", op.buf);
      g_drained[op.buf] = false;
      g_fresh[op.buf] = true;
      acc.clear();
    } else if (acc.empty()) {
      fail("accumulating into an accumulator that was never started", op.buf);
    }
    acc.push_back(tile);
  } else if (!op.accumulate || acc.empty() || acc.back() != tile || g_acc_ks[op.buf] != op.ks) {
    fail("k-steps of a tile out of order", tile, op.ks, g_acc_ks[op.buf]);
  }
  g_acc_ks[op.buf] = (op.ks + 1) % 8;
}

// ---- the kernel's role code, one thread per warp -------------------------------------------------------------
static W4Params g_p;
static int g_u_begin, g_u_end;

// a template like the kernel, so that its `if constexpr` branches are discarded the same way
template <int MT, int NSUB, bool TRACE, int VAR>
static void role_main(int warp_id) {
  using Cfg = W4Cfg<MT, NSUB>;
  threadIdx.x = warp_id * 32;
  blockIdx.x = 0;
  const int warp = warp_id, lane = 0;
  const W4Params& p = g_p;
  const int KT = p.KT;
  const int u_begin = g_u_begin, u_end = g_u_end;
  const uint32_t tmem_base = 0;
  CUtensorMap amap{0};
  (void)lane;
#include "w4_emu_roles.inc"
}

static std::atomic<bool> g_stop{false};
static void hardware_main(uint32_t seed) {  // copies complete in any order, the pipe executes in order
  std::mt19937 rng(seed);
  long held_tile = -1;
  int held_for = 0;
  while (!g_stop.load()) {
    {
      std::lock_guard<std::mutex> lk(g_mu);
      const int what = (int)(rng() % 3);
      // adversarial memory system: now and then one copy is held back for a long while (a late
      // blob) while everything else makes progress — the ring rules must hold regardless
      if (held_for == 0 && !g_copies.empty() && rng() % 64 == 0) {
        held_tile = g_copies[rng() % g_copies.size()].tile;
        held_for = 400 + (int)(rng() % 3000);
      }
      if (held_for > 0) --held_for;
      if (what == 0 && !g_copies.empty()) {
        size_t i = rng() % g_copies.size();
        if (held_for > 0 && g_copies[i].kind == 0 && g_copies[i].tile == held_tile) {
          if (g_copies.size() == 1) continue;       // only the held copy is in flight: let time pass
          i = (i + 1) % g_copies.size();
          if (g_copies[i].kind == 0 && g_copies[i].tile == held_tile) continue;
        }
        const Copy c = g_copies[i];
        g_copies.erase(g_copies.begin() + (long)i);
        if (c.kind == 0) g_raw_content[c.stage] = c.tile;
        else g_act_content[c.stage][c.half] = c.tile;
        bar_complete_tx(c.bar, c.bytes);
      } else if (what == 1 && !g_pipe.empty()) {
        const PipeOp op = g_pipe.front();
        g_pipe.pop_front();
        pipe_execute(op);
      }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
}

static int run_cta(int KT, int u_begin, int u_end, uint32_t seed) {
  g_failed = false;
  g_stop = false;
  g_copies.clear();
  g_pipe.clear();
  g_segments.clear();
  g_seg_committed = 0;
  g_KT = KT;
  for (auto& a : g_acc) a.clear();
  g_acc_ks[0] = g_acc_ks[1] = 0;
  g_drained[0] = g_drained[1] = true;
  g_fresh[0] = g_fresh[1] = false;
  g_readable[0] = g_readable[1] = false;
  g_reads[0] = g_reads[1] = 0;
  for (auto& x : g_raw_content) x = -1;
  for (auto& s : g_act_content) s[0] = s[1] = -1;
  for (auto& s : g_slot_content)
    for (auto& q : s) q[0] = q[1] = -1;
  for (int u = u_begin; u < u_end;) {  // the CTA's accumulator segments, as SegIter walks them
    const int nt = u / KT, kt0 = u - nt * KT, kt1 = std::min(KT, kt0 + (u_end - u));
    g_segments.push_back(Segment{nt, kt0, kt1});
    u += kt1 - kt0;
  }
  g_p = W4Params{};
  g_p.packed = PACKED_BASE;
  g_p.partials = nullptr;
  g_p.M = 0;  // no partial stores: the epilogue only runs its hand-shakes
  g_p.N = 128 * 64;
  g_p.KT = KT;
  g_p.geff_log2 = 7;
  g_p.ngrp = 1;
  g_p.blob_bytes = 8192 + 256 + 128;
  g_p.plan = W4Plan{KT * 64, 1, KT, 64, 1, 0};
  g_u_begin = u_begin;
  g_u_end = u_end;
  {
    // the kernel's own barrier initialisation
#include "w4_emu_init.inc"
  }
  std::thread hw(hardware_main, seed);
  std::vector<std::thread> warps;
  for (int w = 0; w < W4_THREADS / 32; ++w) warps.emplace_back(role_main<MT, NSUB, TRACE, VAR>, w);
  std::atomic<int> joined{0};
  std::thread watchdog([&] {
    for (int i = 0; i < 30000 && joined.load() == 0; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    if (joined.load() == 0) fail("deadlock: the roles did not finish", u_begin, u_end, KT);
  });
  for (auto& t : warps) t.join();
  joined = 1;
  watchdog.join();
  // drain what is still queued, then stop the hardware
  for (int i = 0; i < 1000; ++i) {
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (g_pipe.empty() && g_copies.empty()) break;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  g_stop = true;
  hw.join();
  if (!g_failed.load() && g_seg_committed != g_segments.size())
    fail("not every segment was committed", (long)g_seg_committed, (long)g_segments.size());
  return g_failed.load() ? 1 : 0;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? std::atoi(argv[1]) : 12;
  std::mt19937 rng(1234 + EMU_VAR);
  int bad = 0, n = 0;
  const int kts[] = {1, 2, 3, 8, 32};
  for (int r = 0; r < rounds && !bad; ++r)
    for (int KT : kts) {
      const int share = 1 + (int)(rng() % (unsigned)(2 * KT + 9));
      const int u_begin = (int)(rng() % (unsigned)(3 * KT));
      bad += run_cta(KT, u_begin, u_begin + share, (uint32_t)rng());
      ++n;
      if (bad) break;
    }
  std::printf("w4_emu VAR=%d MT=%d: %d CTA runs, %s\n", EMU_VAR, EMU_MT, n, bad ? "FAILED" : "ok");
  return bad ? 1 : 0;
}
