// readbw.cu — what HBM read bandwidth does each load path sustain on B200?
//   mode 0: LDG.128 (ld.global.nc.L1::no_allocate), U loads in flight per thread
//   mode 1: per-warp ring of cp.async.bulk (1-D) copies of CHUNK bytes, STAGES deep
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o readbw readbw.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int U>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ src, size_t n_vec,
                                                  uint32_t* out) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n_vec; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                   : "l"(src + i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// each warp streams its own contiguous region with a ring of bulk copies
template <int STAGES>
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ src, size_t bytes,
                                                   int chunk, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  uint8_t* my = smem + (size_t)warp * STAGES * chunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)W * STAGES * chunk) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const size_t n_warps = (size_t)gridDim.x * W;
  const size_t gw = (size_t)blockIdx.x * W + warp;
  const size_t n_chunks = bytes / chunk;
  // chunk c of warp gw: index gw + c * n_warps  (interleaved so all warps sweep the buffer together)
  const size_t my_n = (n_chunks > gw) ? (n_chunks - gw + n_warps - 1) / n_warps : 0;
  auto issue = [&](size_t c) {
    const int s = c % STAGES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])),
                 "r"(chunk)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(my + (size_t)s * chunk)),
        "l"(src + (gw + c * n_warps) * (size_t)chunk), "r"(chunk), "r"(smem_u32(&bars[s]))
        : "memory");
  };
  if (lane == 0)
    for (size_t c = 0; c < STAGES && c < my_n; ++c) issue(c);
  uint32_t acc = 0;
  for (size_t c = 0; c < my_n; ++c) {
    const int s = c % STAGES;
    const uint32_t ph = (c / STAGES) & 1;
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bars[s])), "r"(ph)
          : "memory");
    acc ^= reinterpret_cast<const uint32_t*>(my + (size_t)s * chunk)[lane];
    __syncwarp();
    if (lane == 0 && c + STAGES < my_n) issue(c + STAGES);
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)4 << 30;  // 4 GiB >> L2
  uint8_t* buf;
  uint32_t* out;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&out, 4);
  cudaMemset(buf, 1, bytes);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  const size_t n_vec = bytes / 16;
  for (int occ : {2, 4, 8}) {
    float ms;
    ms = time_ms([&] { ldg_kernel<4><<<sms * occ, 256>>>((const uint4*)buf, n_vec, out); }, 3);
    printf("LDG.128 U=4  blocks/SM=%d : %7.1f GB/s\n", occ, bytes / ms / 1e6);
    ms = time_ms([&] { ldg_kernel<8><<<sms * occ, 256>>>((const uint4*)buf, n_vec, out); }, 3);
    printf("LDG.128 U=8  blocks/SM=%d : %7.1f GB/s\n", occ, bytes / ms / 1e6);
    ms = time_ms([&] { ldg_kernel<16><<<sms * occ, 256>>>((const uint4*)buf, n_vec, out); }, 3);
    printf("LDG.128 U=16 blocks/SM=%d : %7.1f GB/s\n", occ, bytes / ms / 1e6);
  }
  for (int chunk : {2048, 4096, 8192, 16384}) {
    for (int ctas : {1, 2}) {
      auto run = [&](auto kernel, int stages) {
        const size_t smem = (size_t)4 * stages * chunk + 4 * stages * 8 + 64;
        if (smem * ctas > 220 * 1024) return;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float ms = time_ms([&] { kernel<<<sms * ctas, 128, smem>>>(buf, bytes, chunk, out); }, 3);
        printf("bulk chunk=%5d stages=%d CTAs/SM=%d (in flight/SM %4zu KB): %7.1f GB/s\n", chunk,
               stages, ctas, (size_t)4 * stages * chunk * ctas / 1024, bytes / ms / 1e6);
      };
      run(bulk_kernel<2>, 2);
      run(bulk_kernel<3>, 3);
      run(bulk_kernel<4>, 4);
      run(bulk_kernel<8>, 8);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
