// kvbw.cu — isolates the memory-side behaviour of the paged-attention access pattern on B200:
// a [n_slots, 8 heads, 128] bf16 cache read tile by tile (16 slots x 256 B of ONE head, blocks of
// 8 slots at random places) through per-warp rings, with different copy engines / shapes:
//   mode 0: tensor TMA 4-D box {64,2,1,8}, SWIZZLE_128B   (what paged_attn_mma_kernel does)
//   mode 1: tensor TMA 3-D box {128,1,8}, no swizzle       (what the CUDA-core kernel does)
//   mode 2: 1-D bulk copies, one per slot row (256 B each)
//   mode 3: tensor TMA 3-D box {128,8,8}: ALL 8 heads of a block in one op (16 KB contiguous)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o kvbw kvbw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(ph)
        : "memory");
}

constexpr int D = 128, HKV = 8, TILE = 16, BS = 8;
constexpr int ROWB = D * 2;

// grid: (splits, heads_per_launch, seqs); 4 warps per CTA interleave tiles like the real kernel
template <int MODE, int STAGES>
__global__ void __launch_bounds__(128) kv_kernel(const __grid_constant__ CUtensorMap map,
                                                 const uint8_t* __restrict__ base,
                                                 const int* __restrict__ table, int blocks_per_seq,
                                                 int tiles_per_split, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int TILE_BYTES = (MODE == 3 ? HKV : 1) * TILE * ROWB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem + (size_t)warp * STAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)4 * STAGES * TILE_BYTES) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int split = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int* tbl = table + (size_t)seq * blocks_per_seq;
  const int t0 = split * tiles_per_split;
  const int n_my = (tiles_per_split - warp + 3) / 4;
  auto issue = [&](int i) {
    const int tile = t0 + warp + i * 4;
    const int s = i % STAGES;
    uint8_t* dst = my + (size_t)s * TILE_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])),
                 "r"(TILE_BYTES)
                 : "memory");
    for (int bx = 0; bx < TILE / BS; ++bx) {
      const int slot0 = tbl[tile * (TILE / BS) + bx];
      if (MODE == 0) {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst + bx * BS * ROWB)),
            "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(0), "r"(head), "r"(slot0)
            : "memory");
      } else if (MODE == 1) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst + bx * BS * ROWB)),
            "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(head), "r"(slot0)
            : "memory");
      } else if (MODE == 2) {
        for (int r = 0; r < BS; ++r)
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(dst + (bx * BS + r) * ROWB)),
              "l"(base + ((size_t)(slot0 + r) * HKV + head) * ROWB), "r"(ROWB), "r"(smem_u32(&bars[s]))
              : "memory");
      } else {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst + bx * BS * HKV * ROWB)),
            "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(0), "r"(slot0)
            : "memory");
      }
    }
  };
  if (lane == 0)
    for (int i = 0; i < STAGES && i < n_my; ++i) issue(i);
  uint32_t acc = 0;
  for (int i = 0; i < n_my; ++i) {
    const int s = i % STAGES;
    mbar_wait(&bars[s], (i / STAGES) & 1);
    acc ^= reinterpret_cast<const uint32_t*>(my + (size_t)s * TILE_BYTES)[lane];
    __syncwarp();
    if (lane == 0 && i + STAGES < n_my) issue(i + STAGES);
  }
  if (acc == 0x12345678u) out[0] = acc;
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill);

int main() {
  const int B = 64, S = 2048, L = 4;  // L independent caches, rotated (each 268 MB)
  const int blocks_per_seq = S / BS, n_blocks = B * blocks_per_seq + 64;
  const size_t n_slots = (size_t)n_blocks * BS, bytes = n_slots * HKV * ROWB;
  std::vector<uint8_t*> caches(L);
  for (auto& c : caches) {
    cudaMalloc(&c, bytes);
    cudaMemset(c, 1, bytes);
  }
  std::vector<int> perm(n_blocks);
  for (int i = 0; i < n_blocks; ++i) perm[i] = i;
  std::mt19937 rng(2);
  std::shuffle(perm.begin(), perm.end(), rng);
  std::vector<int> table(B * blocks_per_seq), table_seq(B * blocks_per_seq);
  for (int i = 0; i < B * blocks_per_seq; ++i) {
    table[i] = perm[i] * BS;
    table_seq[i] = i * BS;
  }
  int *d_table, *d_table_seq;
  uint32_t* out;
  cudaMalloc(&d_table, table.size() * 4);
  cudaMalloc(&d_table_seq, table.size() * 4);
  cudaMalloc(&out, 4);
  cudaMemcpy(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_table_seq, table_seq.data(), table.size() * 4, cudaMemcpyHostToDevice);

  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)fp;
  auto make_map = [&](int mode, uint8_t* ptr) {
    CUtensorMap m;
    CUresult r;
    if (mode == 0) {
      cuuint64_t dims[4] = {64, 2, HKV, n_slots};
      cuuint64_t str[3] = {128, ROWB, (cuuint64_t)HKV * ROWB};
      cuuint32_t box[4] = {64, 2, 1, BS}, es[4] = {1, 1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, str, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[3] = {D, HKV, n_slots};
      cuuint64_t str[2] = {ROWB, (cuuint64_t)HKV * ROWB};
      cuuint32_t box[3] = {D, (cuuint32_t)(mode == 3 ? HKV : 1), BS}, es[3] = {1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, str, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
  };

  auto bench = [&](auto kernel, int mode, int stages, int splits, const int* tbl, const char* tag) {
    const int tile_bytes = (mode == 3 ? HKV : 1) * TILE * ROWB;
    const size_t smem = (size_t)4 * stages * tile_bytes + 4 * stages * 8 + 64;
    if (smem > 200 * 1024) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<CUtensorMap> maps;
    for (auto c : caches) maps.push_back(make_map(mode, c));
    const int tiles = S / TILE, tps = tiles / splits;
    dim3 grid(splits, mode == 3 ? 1 : HKV, B);
    auto launch = [&](int l) {
      kernel<<<grid, 128, smem>>>(maps[l], caches[l], tbl, blocks_per_seq, tps, out);
    };
    for (int l = 0; l < L; ++l) launch(l);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; ++r)
      for (int l = 0; l < L; ++l) launch(l);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double per = ms / (reps * L);
    const double byts = (double)B * S * HKV * ROWB;  // one of K/V
    printf("mode %d %-22s stages=%d splits=%2d smem=%3zu KB: %7.1f us  %7.1f GB/s  (%s)\n", mode, tag,
           stages, splits, smem / 1024, per * 1e3, byts / per / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int splits : {4, 8}) {
    bench(kv_kernel<0, 3>, 0, 3, splits, d_table, "tma4d swz random");
    bench(kv_kernel<0, 6>, 0, 6, splits, d_table, "tma4d swz random");
    bench(kv_kernel<0, 6>, 0, 6, splits, d_table_seq, "tma4d swz sequential");
    bench(kv_kernel<1, 6>, 1, 6, splits, d_table, "tma3d random");
    bench(kv_kernel<2, 6>, 2, 6, splits, d_table, "bulk1d rows random");
    bench(kv_kernel<3, 2>, 3, 2, splits, d_table, "tma3d all-heads random");
    bench(kv_kernel<3, 3>, 3, 3, splits, d_table, "tma3d all-heads random");
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
