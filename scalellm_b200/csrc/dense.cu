// dense.cu — dense bf16 linear for decode-sized batches: C[M, N] = A[M, K] W[N, K]^T
// (+ bias), M <= 128 rows per pass, replacing F::linear -> cuBLASLt
// (src/layers/linear/parallel_linear.cpp:256-263,294-308; lm_head models/meta/llama.h:259-265) for the
// unquantised configuration (BASELINE config 2) and for lm_head.
//
// HBM bound (AI = M FLOP/B <= 128 against a ridge of ~280): the job is to stream W once at HBM rate.
// Same skeleton as the W4A16 kernel minus the dequant stage — "swap AB": the weight tile is the MMA's
// A operand (M = 128 output features), the activations its B operand (N = tokens):
//   * persistent, one CTA per SM, stream-K over (n tile of 128, k tile of 128) units (common.cuh
//     W4Plan: the fp32 partials and their consumers are the W4A16 ones);
//   * warp 0: weight producer — TMA 2-D {64 k, 128 n} boxes, SWIZZLE_128B, two per k tile, into a
//     ring of 32 KB stages; weights are constants, so it does not execute griddepcontrol.wait and
//     streams while the predecessor kernel still runs;
//   * warp 1: activation producer — griddepcontrol.wait, then TMA {64 k, MT tokens} boxes;
//   * warp 2: MMA issuer — tcgen05.mma.cta_group::1.kind::f16, A and B from shared memory
//     (K-major, 128-byte swizzle), D [128 n x MT tokens] fp32 in TMEM, double buffered;
//     tcgen05.commit releases the stages;
//   * warps 4-7: epilogue — tcgen05.ld -> fp32 partial of the segment (lane <-> n).
// Rows of W past N and columns past K are zero-filled by TMA (out-of-bounds fill), so N and K need
// only be multiples of 8 resp. 64: vocabulary shards like 16 032 columns work.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

template <int MT>
struct DenseCfg {
  static constexpr int W_ATOM = 128 * 128;            // bytes of one [128 n x 64 k] bf16 swizzle atom
  static constexpr int W_BYTES = 2 * W_ATOM;          // 128 k per stage: 32 KB
  static constexpr int A_ATOM = MT * 128;             // [MT x 64 k]
  static constexpr int A_BYTES = 2 * A_ATOM;
  static constexpr int STAGES = MT <= 64 ? 4 : 3;     // weights + activations share a stage index
  static constexpr int TMEM_COLS = 2 * MT < 32 ? 32 : 2 * MT;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * (W_BYTES + A_BYTES) + (3 * STAGES + 4) * 8 + 64;
};

constexpr int DN_THREADS = 8 * 32;
constexpr int DN_WARP_W = 0, DN_WARP_A = 1, DN_WARP_MMA = 2, DN_WARP_EPI = 4;

struct DenseParams {
  float* partials;      // [slots][M][N] fp32
  int64_t slot_stride;  // M * N
  int M, N, KT;
  W4Plan plan;
};

struct DnSeg {
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

template <int MT>
__global__ void __launch_bounds__(DN_THREADS, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap amap,
                  const DenseParams p) {
  using Cfg = DenseCfg<MT>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* w_smem = base;
  uint8_t* a_smem = w_smem + Cfg::STAGES * Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + Cfg::STAGES * Cfg::A_BYTES);
  uint64_t* w_full = bars;
  uint64_t* a_full = w_full + Cfg::STAGES;
  uint64_t* empty = a_full + Cfg::STAGES;   // released by the MMA's commit (both operands)
  uint64_t* tmem_full = empty + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KT = p.KT;
  const int u_begin = w4_unit_begin(blockIdx.x, p.plan.units, p.plan.P);
  const int u_end = w4_unit_begin(blockIdx.x + 1, p.plan.units, p.plan.P);

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&a_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == DN_WARP_MMA) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp == DN_WARP_W && lane == 0) prefetch_tensormap(&wmap);
  if (warp == DN_WARP_A && lane == 0) prefetch_tensormap(&amap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == DN_WARP_W) {
    if (lane == 0) {  // weights are never written by another kernel: no griddepcontrol.wait
      DnSeg it{u_begin, u_end, KT};
      int nt, kt0, kt1, cnt = 0;
      while (it.next(nt, kt0, kt1)) {
        for (int kt = kt0; kt < kt1; ++kt, ++cnt) {
          const int s = cnt % Cfg::STAGES;
          const uint32_t ph = (cnt / Cfg::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&w_full[s], (uint32_t)Cfg::W_BYTES);
          uint8_t* dst = w_smem + s * Cfg::W_BYTES;
          tma_load_2d(dst, &wmap, &w_full[s], kt * 128, nt * 128);
          tma_load_2d(dst + Cfg::W_ATOM, &wmap, &w_full[s], kt * 128 + 64, nt * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == DN_WARP_A) {
    if (lane == 0) {
      pdl_wait();  // activations are the predecessor kernel's output
      DnSeg it{u_begin, u_end, KT};
      int nt, kt0, kt1, cnt = 0;
      while (it.next(nt, kt0, kt1)) {
        for (int kt = kt0; kt < kt1; ++kt, ++cnt) {
          const int s = cnt % Cfg::STAGES;
          const uint32_t ph = (cnt / Cfg::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&a_full[s], (uint32_t)Cfg::A_BYTES);
          uint8_t* dst = a_smem + s * Cfg::A_BYTES;
          tma_load_2d(dst, &amap, &a_full[s], kt * 128, 0);
          tma_load_2d(dst + Cfg::A_ATOM, &amap, &a_full[s], kt * 128 + 64, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == DN_WARP_MMA) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, MT);
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t w_base = __shfl_sync(0xffffffffu, smem_u32(w_smem), 0);
    const uint32_t a_base = __shfl_sync(0xffffffffu, smem_u32(a_smem), 0);
    DnSeg it{u_begin, u_end, KT};
    int nt, kt0, kt1, cnt = 0, seg = 0;
    while (it.next(nt, kt0, kt1)) {
      const int buf = seg & 1;
      mbar_wait(&tmem_empty[buf], ((seg >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tbase + buf * MT;
      for (int kt = kt0; kt < kt1; ++kt, ++cnt) {
        const int s = cnt % Cfg::STAGES;
        const uint32_t ph = (cnt / Cfg::STAGES) & 1;
        mbar_wait(&w_full[s], ph);
        mbar_wait(&a_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wd0 = umma_desc_kmajor_sw128(w_base + s * Cfg::W_BYTES);
          const uint64_t ad0 = umma_desc_kmajor_sw128(a_base + s * Cfg::A_BYTES);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            // descriptor start-address field is in 16-byte units: 32 bytes per 16-k step inside an
            // atom, the second atom (k 64..127) one atom size further
            const uint64_t wd = wd0 + (uint64_t)(((ks >> 2) * Cfg::W_ATOM + (ks & 3) * 32) >> 4);
            const uint64_t ad = ad0 + (uint64_t)(((ks >> 2) * Cfg::A_ATOM + (ks & 3) * 32) >> 4);
            umma_bf16(d_tmem, wd, ad, idesc, (ks > 0 || kt > kt0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // the stage's 2 producers each wait on it
          if (kt == kt1 - 1) umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
      }
      ++seg;
    }
    __syncwarp();
  } else if (warp >= DN_WARP_EPI) {
    const int quad = warp & 3;
    const int n_local = quad * 32 + lane;
    pdl_wait();  // the partials buffer may still be read by an earlier kernel's consumer
    DnSeg it{u_begin, u_end, KT};
    int nt, kt0, kt1, seg = 0;
    while (it.next(nt, kt0, kt1)) {
      const int buf = seg & 1;
      const int slot = (int)blockIdx.x - w4_first_owner(p.plan, nt);
      const int n = nt * 128 + n_local;
      float* part = p.partials + (int64_t)slot * p.slot_stride + n;
      mbar_wait(&tmem_full[buf], (seg >> 1) & 1);
      if (it.u >= it.u1 && warp == DN_WARP_EPI && lane == 0) pdl_launch_dependents();
      tc_fence_after();
      constexpr int CH = MT >= 32 ? 32 : 16;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * MT;
#pragma unroll 1
      for (int c0 = 0; c0 < MT; c0 += CH) {
        uint32_t r[CH];
        if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, r);
        else tmem_ld_32x32b_x16(taddr + c0, r);
        tmem_ld_wait();
        if (n < p.N) {
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int m = c0 + i;
            if (m < p.M) part[(int64_t)m * p.N] = __uint_as_float(r[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DN_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct DMapKey {
  const void* ptr;
  int64_t rows, K, ld;
  int box_rows;
  bool operator==(const DMapKey& o) const {
    return ptr == o.ptr && rows == o.rows && K == o.K && ld == o.ld && box_rows == o.box_rows;
  }
};
static std::mutex g_dmap_mu;
static std::vector<std::pair<DMapKey, CUtensorMap>> g_dmaps;

// [rows, K] bf16 row-major (row stride ld elements) as K-major SWIZZLE_128B boxes {64 k, box_rows}
static int get_kmajor_map(const DMapKey& key, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_dmap_mu);
    for (const auto& e : g_dmaps)
      if (e.first == key) {
        *out = e.second;
        return B200_OK;
      }
  }
  tensor_map_encode_fn enc = get_tensor_map_encode();
  if (!enc) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)key.K, (cuuint64_t)key.rows};
  cuuint64_t strides[1] = {(cuuint64_t)key.ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)key.box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(key.ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a dense operand", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_dmap_mu);
    if (g_dmaps.size() > 4096) g_dmaps.clear();
    g_dmaps.push_back({key, m});
  }
  *out = m;
  return B200_OK;
}

static int dn_pick_mt(int64_t M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128; }
static int64_t up128(int64_t x) { return (x + 127) / 128 * 128; }

template <int MT>
static int launch_dense(const CUtensorMap& wmap, const CUtensorMap& amap, const DenseParams& p,
                        cudaStream_t st) {
  using Cfg = DenseCfg<MT>;
  auto kern = dense_gemm_kernel<MT>;
  B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  B200_PDL_LAUNCH_L(1, "dense_gemm", kern, (unsigned)p.plan.P, DN_THREADS, Cfg::SMEM, st, wmap, amap, p);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_dense_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 256;
  const int64_t mc = M < 128 ? M : 128;
  const W4Plan plan = w4_get_plan(up128(N), up128(K), 128);
  return (int64_t)plan.slots * mc * N * (int64_t)sizeof(float) + 256;
}

int b200_dense_gemm(void* C, const void* A, const void* W, const void* bias, int64_t M, int64_t N,
                    int64_t K, int64_t lda, int64_t ldw, int64_t ldc, void* workspace,
                    int64_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(C && A && W, "dense_gemm: null pointer");
  B200_CHECK_ARG(M >= 0 && N > 0 && K > 0 && N % 8 == 0 && K % 64 == 0,
                 "dense_gemm: N %% 8 == 0 and K %% 64 == 0 required (N=%lld, K=%lld)", (long long)N, (long long)K);
  B200_CHECK_ARG(lda >= K && ldw >= K && ldc >= N && lda % 8 == 0 && ldw % 8 == 0 && is_aligned(A, 16) &&
                     is_aligned(W, 16) && is_aligned(C, 2),
                 "dense_gemm: rows of A and W must be 16-byte aligned and dense in K");
  if (M == 0) return B200_OK;
  B200_CHECK_ARG(workspace && is_aligned(workspace, 16), "dense_gemm: workspace required");
  auto st = static_cast<cudaStream_t>(stream);
  const W4Plan plan = w4_get_plan(up128(N), up128(K), 128);   // one weight tile per unit
  CUtensorMap wmap;
  int rc = get_kmajor_map(DMapKey{W, N, K, ldw, 128}, &wmap);
  if (rc != B200_OK) return rc;
  for (int64_t m0 = 0; m0 < M; m0 += 128) {   // larger batches: one pass over W per 128 rows
    const int64_t mc = (M - m0) < 128 ? (M - m0) : 128;
    const int64_t need = (int64_t)plan.slots * mc * N * (int64_t)sizeof(float);
    if (workspace_bytes < need)
      return set_error(B200_ERR_WORKSPACE, "dense_gemm: workspace %lld B < required %lld B",
                       (long long)workspace_bytes, (long long)need);
    const int mt = dn_pick_mt(mc);
    const __nv_bfloat16* a0 = static_cast<const __nv_bfloat16*>(A) + m0 * lda;
    CUtensorMap amap;
    rc = get_kmajor_map(DMapKey{a0, mc, K, lda, mt}, &amap);
    if (rc != B200_OK) return rc;
    DenseParams p{};
    p.partials = static_cast<float*>(workspace);
    p.slot_stride = mc * N;
    p.M = (int)mc;
    p.N = (int)N;
    p.KT = plan.KT;
    p.plan = plan;
    switch (mt) {
      case 16: rc = launch_dense<16>(wmap, amap, p, st); break;
      case 32: rc = launch_dense<32>(wmap, amap, p, st); break;
      case 64: rc = launch_dense<64>(wmap, amap, p, st); break;
      default: rc = launch_dense<128>(wmap, amap, p, st); break;
    }
    if (rc != B200_OK) return rc;
    rc = w4_launch_reduce(static_cast<__nv_bfloat16*>(C) + m0 * ldc, p.partials,
                          static_cast<const __nv_bfloat16*>(bias), (int)mc, (int)N, ldc, mc * N, plan, st);
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}

int b200_dense_splitk_splits(int64_t M, int64_t N, int64_t K) {
  (void)M;
  if (N <= 0 || K <= 0 || N % 128 || K % 128) return 0;
  return w4_get_plan(N, K, 128).slots;
}

int b200_dense_gemm_splitk(float* partials, const void* A, const void* W, int64_t M, int64_t N,
                           int64_t K, int64_t lda, int64_t ldw, int splits, b200_stream_t stream) {
  B200_CHECK_ARG(partials && A && W, "dense_gemm_splitk: null pointer");
  B200_CHECK_ARG(M > 0 && M <= 128 && N > 0 && K > 0 && N % 128 == 0 && K % 128 == 0,
                 "dense_gemm_splitk: 1 <= M <= 128, N %% 128 == 0, K %% 128 == 0 required");
  B200_CHECK_ARG(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0 && is_aligned(A, 16) &&
                     is_aligned(W, 16) && is_aligned(partials, 16),
                 "dense_gemm_splitk: rows of A and W must be 16-byte aligned and dense in K");
  const W4Plan plan = w4_get_plan(N, K, 128);
  B200_CHECK_ARG(splits == plan.slots, "dense_gemm_splitk: partials must have b200_dense_splitk_splits() = %d slots, got %d",
                 plan.slots, splits);
  auto st = static_cast<cudaStream_t>(stream);
  CUtensorMap wmap, amap;
  int rc = get_kmajor_map(DMapKey{W, N, K, ldw, 128}, &wmap);
  if (rc != B200_OK) return rc;
  const int mt = dn_pick_mt(M);
  rc = get_kmajor_map(DMapKey{A, M, K, lda, mt}, &amap);
  if (rc != B200_OK) return rc;
  DenseParams p{};
  p.partials = partials;
  p.slot_stride = M * N;
  p.M = (int)M;
  p.N = (int)N;
  p.KT = plan.KT;
  p.plan = plan;
  switch (mt) {
    case 16: return launch_dense<16>(wmap, amap, p, st);
    case 32: return launch_dense<32>(wmap, amap, p, st);
    case 64: return launch_dense<64>(wmap, amap, p, st);
    default: return launch_dense<128>(wmap, amap, p, st);
  }
}

}  // extern "C"
