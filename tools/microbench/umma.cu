// umma.cu — tcgen05.mma (kind::f16, bf16, M=128, K=16) throughput on one SM as a function of N and
// of where A comes from: tensor memory (.ts form, what the W4A16 GEMM uses for dequantised weights)
// or shared memory (descriptor).  One CTA per SM, one thread issues GROUPS x 8 MMAs, one commit per
// group of 8 (= one 128-k weight tile), then waits for the last commit.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../scalellm_b200/csrc/common.cuh"
using namespace b200;

template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k_umma(int groups, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t holder;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 256) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // small bf16 values
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = holder;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  if (warp == 0) {
    const uint64_t a_desc0 = umma_desc_kmajor_sw128(smem_u32(sm));            // [128 x 64] bf16 atom
    const uint64_t b_desc0 = umma_desc_kmajor_sw128(smem_u32(sm + 16384));   // [N x 64] bf16 atom
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t b_desc = b_desc0 + (uint64_t)(((ks & 3) * 32) >> 4);
          if (A_TMEM)
            umma_bf16_ts(tbase + (g & 1) * N, tbase + 256 + (g & 3) * 64 + ks * 8, b_desc, idesc, 1u);
          else
            umma_bf16(tbase + (g & 1) * N, a_desc0 + (uint64_t)(((ks & 3) * 32) >> 4), b_desc, idesc, 1u);
        }
        if (g == groups - 1) umma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int N, bool A_TMEM>
void run(int sms) {
  long long* cyc;
  cudaMalloc(&cyc, 8 * sms);
  const size_t smem = 16384 + N * 256 + 1024;
  cudaFuncSetAttribute(k_umma<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int groups = 2000;
  k_umma<N, A_TMEM><<<sms, 128, smem>>>(groups, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("M=128 N=%3d K=16 A from %-6s: %6.1f cycles per MMA, %7.1f per 8 (one 128-k tile)  [%s]\n", N,
         A_TMEM ? "TMEM" : "smem", (double)h / groups / 8, (double)h / groups, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  run<64, true>(sms);
  run<64, false>(sms);
  run<128, true>(sms);
  run<128, false>(sms);
  run<32, true>(sms);
  run<32, false>(sms);
  run<16, true>(sms);
  return 0;
}
