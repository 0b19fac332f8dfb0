// hmma.cu — latency / throughput of legacy mma.sync.m16n8k16 (bf16) and ldmatrix on B200.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ILP independent accumulator chains per warp, N iterations
template <int ILP>
__global__ void k_mma(int iters, long long* cyc, float* out) {
  uint32_t a[4] = {threadIdx.x, 2, 3, 4};
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma(d[i], a, 0x3f803f80u + i, 0x3f803f80u);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (s == 12345.f) out[0] = s;
}

__global__ void k_ldsm(int iters, long long* cyc, uint32_t* out) {
  __shared__ __align__(1024) uint8_t sm[16384];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  __syncthreads();
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r0, r1, r2, r3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr + (acc & 0x3) * 512));
    acc += r0 ^ r1 ^ r2 ^ r3;   // dependent: measures latency
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 12345) out[0] = acc;
}

int main() {
  long long* cyc; float* out;
  cudaMalloc(&cyc, 8); cudaMalloc(&out, 4);
  const int iters = 2000;
  auto rd = [&] { long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); return (double)h; };
  // latency: 1 warp, 1 chain
  k_mma<1><<<1, 32>>>(iters, cyc, out); cudaDeviceSynchronize();
  printf("mma.m16n8k16 bf16 dependent-chain latency: %.1f cycles\n", rd() / iters);
  for (int warps : {1, 2, 4, 8, 16}) {
    k_mma<8><<<1, 32 * warps>>>(iters, cyc, out); cudaDeviceSynchronize();
    double c = rd();
    printf("mma ILP=8 warps/SM=%2d: %.2f cycles per MMA per warp, SM throughput %.1f MMA/100cyc = %.0f MAC/clk/SM\n",
           warps, c / (iters * 8), 100.0 * iters * 8 * warps / c, 2048.0 * iters * 8 * warps / c);
  }
  k_ldsm<<<1, 32>>>(iters, cyc, (uint32_t*)out); cudaDeviceSynchronize();
  printf("ldmatrix.x4 dependent latency: %.1f cycles\n", rd() / iters);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
