// Known-answer generator for the Marlin int4 -> bf16 weight arithmetic, built from the reference's
// OWN device functions (src/kernels/quantization/marlin/numeric_conversion.h: dequant<> :144-167,
// scale :221-229, sub_zp :232-240), applied in the order gemm_kernel.cuh:715-765 applies them.
// TEST INFRASTRUCTURE: pins oracle/quant.py's `bf16_mul(bf16(q) - bf16(z), s)` restatement (and
// through it b200_w4a16_dequant) to the reference code for every (q, z) pair — the AWQ zero-point
// path has no known-answer test in the reference itself (tests/kernels/marlin_gemm_test.py:97).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "numeric_conversion.h"

namespace {

using bf16 = nv_bfloat16;

// out_zp[q][z][s] : has_zp = true  (AWQ, GPTQ with zero points)
// out_sym[q][s]   : has_zp = false (symmetric GPTQ: the -8 is folded into dequant)
__global__ void marlin_dequant_kat_kernel(const bf16* __restrict__ scales, int S,
                                          bf16* __restrict__ out_zp, bf16* __restrict__ out_sym) {
  using namespace marlin;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 16 * 16 * S) return;
  const int s = idx % S, z = (idx / S) % 16, q = idx / (S * 16);
  ScalarType<bf16>::FragS fs;
  fs[0] = ScalarType<bf16>::num2num2(scales[s]);
  {
    ScalarType<bf16>::FragB b = dequant<bf16, 4, true>(q);    // nibble 0 carries q
    ScalarType<bf16>::FragB zf = dequant<bf16, 4, true>(z);   // zero points take the same path (:715)
    nv_bfloat162 zp2 = zf[0];
    sub_zp<bf16>(b, zp2, 0);
    scale<bf16>(b, fs, 0);
    out_zp[idx] = b[0].x;
  }
  if (z == 0) {
    ScalarType<bf16>::FragB b = dequant<bf16, 4, false>(q);
    scale<bf16>(b, fs, 0);
    out_sym[q * S + s] = b[0].x;
  }
}

}  // namespace

extern "C" int marlin_dequant_kat(const void* scales, int S, void* out_zp, void* out_sym,
                                  void* stream) {
  const int total = 16 * 16 * S;
  marlin_dequant_kat_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(scales), S, static_cast<bf16*>(out_zp), static_cast<bf16*>(out_sym));
  return (int)cudaGetLastError();
}
