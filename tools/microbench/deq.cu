// deq.cu — what bounds the W4 dequant warps?  (int4 -> bf16 math, tcgen05.st into TMEM)
// One CTA per SM; WARPS dequant warps (groups of 4 = one TMEM lane quadrant each) loop over
// tiles of 128 n x 128 k exactly like w4a16_gemm_kernel's dequant role.
//   mode 0: math + tcgen05.st     mode 1: math only     mode 2: st only    mode 3: math w/o hsub2
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../scalellm_b200/csrc/common.cuh"
using namespace b200;

__device__ __forceinline__ uint4 dq_word(uint32_t q, __nv_bfloat162 zmagic, __nv_bfloat162 s2, bool sub) {
  uint32_t r[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    uint32_t v;
    asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(v) : "r"(q >> (4 * jj)));
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    if (sub) b = __hsub2(b, zmagic);
    b = __hmul2(b, s2);
    r[jj] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_deq(int tiles, long long* cyc, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 9728 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm)[i] = i * 2654435761u;
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = holder;
  const int group = warp >> 2, n_local = (warp & 3) * 32 + lane;
  const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + group * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < tiles; ++it) {
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
      if (MODE != 2) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 uu = u[hh * 2 + q];
          const uint32_t words[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint4 d = dq_word(words[w], zm, s2, MODE != 3);
            r[(q * 4 + w) * 4 + 0] = d.x; r[(q * 4 + w) * 4 + 1] = d.y;
            r[(q * 4 + w) * 4 + 2] = d.z; r[(q * 4 + w) * 4 + 3] = d.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = u[(i >> 2) & 3].x + i;
      }
      if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      } else {
        if (hh == 0 && it > 0) tmem_st_wait();
        tmem_st_32x32b_x32(lane_base + (it & 1) * 64 + hh * 32, r);
      }
    }
  }
  if (MODE == 0 || MODE == 2) tmem_st_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int MODE>
void run(const char* name, int sms) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 8 * sms); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k_deq<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9728);
  const int tiles = 400;
  for (int warps : {4, 8, 12, 16}) {
    k_deq<MODE><<<sms, warps * 32, 8 * 9728>>>(tiles, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d: %7.1f cycles per 128x128 tile (CTA rate; %s)\n", name, warps,
           (double)h / tiles / (warps / 4), cudaGetErrorString(e));
  }
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  run<0>("math + tcgen05.st", sms);
  run<1>("math only", sms);
  run<2>("tcgen05.st only", sms);
  run<3>("math w/o hsub2", sms);
  return 0;
}
