// mix.cu — does the dequant side (int4->bf16 math + tcgen05.st into the TMEM A ring) slow the
// tensor pipe down?  One CTA per SM: warp 16 issues tcgen05.mma (M=128, N=64, K=16, A from TMEM,
// 8 per "tile", one commit per tile, at most DEPTH tiles in flight) while warps 0..W-1 run the
// W4A16 dequant loop into the same TMEM columns.  Reports cycles per tile for both sides.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../scalellm_b200/csrc/common.cuh"
using namespace b200;

__device__ __forceinline__ uint4 dq_word(uint32_t q, __nv_bfloat162 zmagic, __nv_bfloat162 s2) {
  uint32_t r[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint32_t v;
    asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(v) : "r"(q >> (4 * jj)));
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    b = __hsub2(b, zmagic);
    b = __hmul2(b, s2);
    r[jj] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

// mode bit 0: dequant warps do the math, bit 1: dequant warps do the tcgen05.st, bit 2: MMA runs,
// bit 3: the MMA thread skips the per-group tcgen05.fence::after_thread_sync
__global__ void __launch_bounds__(544, 1) k_mix(int tiles, int mode, int deq_warps, int depth, int every, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t holder;
  __shared__ __align__(8) uint64_t bars[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (8 * 9728 + 16384) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + (i & 0xf0f);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = holder;
  const long long t0 = clock64();
  if (warp < deq_warps && (mode & 3)) {
    const int group = warp >> 2, n_local = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 128 + group * 64;
    uint32_t acc = 0;
    for (int it = 0; it < tiles / (deq_warps / 4); ++it) {
      const uint8_t* raw = sm + (it & 7) * 9728;
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(raw + (q * 128 + n_local) * 16);
      const __nv_bfloat16 sv = reinterpret_cast<const __nv_bfloat16*>(raw + 8192)[n_local];
      const __nv_bfloat162 s2 = __halves2bfloat162(sv, sv);
      const uint32_t m = 0x4300u | (raw[8192 + 256 + n_local] & 15), mm = m | (m << 16);
      const __nv_bfloat162 zm = *reinterpret_cast<const __nv_bfloat162*>(&mm);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32];
        if (mode & 1) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint4 uu = u[hh * 2 + q];
            const uint32_t words[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint4 d = dq_word(words[w], zm, s2);
              r[(q * 4 + w) * 4 + 0] = d.x; r[(q * 4 + w) * 4 + 1] = d.y;
              r[(q * 4 + w) * 4 + 2] = d.z; r[(q * 4 + w) * 4 + 3] = d.w;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = u[(i >> 2) & 3].x + i;
        }
        if (mode & 2) {
          tmem_st_32x32b_x32(lane_base + hh * 32, r);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r[i];
        }
      }
      if (mode & 2) tmem_st_wait();
    }
    if (acc == 0x12345678u) cyc[0] = acc;
    if (warp == 0 && lane == 0) cyc[gridDim.x + blockIdx.x] = clock64() - t0;
  } else if (warp == 16 && (mode & 4)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    // One commit per `every` tiles ("group"); at most `depth` groups in flight before the issuing
    // thread waits for the oldest commit; depth 0 = never wait (single commit at the end);
    // depth < 0 = commit every group but never wait (|depth| barriers used round-robin).
    const uint64_t b_desc0 = umma_desc_kmajor_sw128(smem_u32(sm + 8 * 9728));
    const int groups = tiles / every;
    const int D = depth > 0 ? depth : (depth < 0 ? -depth : 1);
    for (int g = 0; g < groups; ++g) {
      if (depth > 0 && g >= D) mbar_wait(&bars[g % D], ((g / D) - 1) & 1);
      if (!(mode & 8)) tc_fence_after();
      if (elect_one()) {
        for (int t = 0; t < every; ++t) {
          const int tile = g * every + t;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_ts(tbase + (tile & 1) * 64, tbase + 128 + (tile & 3) * 64 + ks * 8,
                         b_desc0 + (uint64_t)(((ks & 3) * 32) >> 4), idesc, 1u);
        }
        if (depth != 0 || g == groups - 1) umma_commit(&bars[depth != 0 ? g % D : 0]);
      }
      __syncwarp();
    }
    if (depth > 0) {
      for (int g = groups > D ? groups - D : 0; g < groups; ++g) mbar_wait(&bars[g % D], (g / D) & 1);
    } else if (depth == 0) {
      mbar_wait(&bars[0], 0);
    } else {
      // commits without waits: the barriers completed many phases; just drain with one more commit
      if (elect_one()) umma_commit(&bars[7]);
      __syncwarp();
      mbar_wait(&bars[7], 0);
    }
    if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  long long* cyc;
  cudaMalloc(&cyc, 16 * sms);
  const size_t smem = 8 * 9728 + 16384 + 1024;
  cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 960;
  struct { int mode, warps, depth, every; const char* name; } cases[] = {
      {4, 16, 0, 1, "MMA alone, no commits"},
      {4 | 8, 16, 0, 1, "MMA alone, no commits, no tcgen05.fence"},
      {4, 16, -4, 1, "MMA alone, commit per tile, never waits"},
      {4 | 8, 16, -4, 1, "MMA alone, commit per tile, no fence"},
      {4, 16, -4, 2, "MMA alone, fence+commit per 2 tiles"},
      {4, 16, 4, 1, "MMA alone, commit+wait per tile (4 deep)"},
      {4, 16, 3, 2, "MMA alone, commit+wait per 2 tiles (3 deep)"},
      {4, 16, 2, 3, "MMA alone, commit+wait per 3 tiles (2 deep)"},
      {4, 16, 2, 6, "MMA alone, commit+wait per 6 tiles (2 deep)"},
      {4 | 3, 16, 0, 1, "MMA no commits + dequant 16 warps"},
      {4 | 3, 16, 4, 1, "MMA commit per tile + dequant 16 warps"},
      {4 | 3, 16, 3, 2, "MMA commit per 2 tiles + dequant 16 warps"},
      {4 | 3, 16, 2, 3, "MMA commit per 3 tiles + dequant 16 warps"},
  };
  for (auto& c : cases) {
    cudaMemset(cyc, 0, 16 * sms);
    k_mix<<<sms, 544, smem>>>(tiles, c.mode, c.warps, c.depth, c.every, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(&h[0], cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&h[1], cyc + sms, 8, cudaMemcpyDeviceToHost);
    printf("%-46s MMA %6.1f cyc/tile   dequant %6.1f cyc/tile  [%s]\n", c.name, (double)h[0] / tiles,
           (double)h[1] / tiles, cudaGetErrorString(e));
  }
  return 0;
}
